// TEST INFRASTRUCTURE: libcipc_shimdrv.so -- the drop-in boundary exercised on the reference's REAL types.
// oracle/ref_build/ref_drivers.cpp (the C API over the reference's six contact templates + five friction templates,
// MESH_NODE / MESH_NODE_ATTR on the real Storage/*.hpp + Math/VECTOR.h) is compiled a second time with
//     -I codim-ipc_b200/shim -I include        in front of        -I oracle/ref_build/stub -I /root/reference/Library
// so that every Compute_* it calls lands in codim-ipc_b200/shim/FEM/IPC.h / FRICTION.h and from there in libcipc_b200.so,
// while the reference's own templates stay available as Compute_*_CPU in the same binary (shim_selfcheck compares the two
// in C++).  Build recipe: tests/shim_harness/Makefile (run by __graft_entry__.build() when /root/reference is present; the
// built library travels to the GPU box).
#define CIPC_SHIM_BUILD 1
#include "../../oracle/ref_build/ref_drivers.cpp"
