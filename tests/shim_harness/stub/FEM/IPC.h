// TEST INFRASTRUCTURE: stand-in for the reference's Library/FEM/IPC.h and the types it drags in, so that
// codim-ipc_b200/shim/FEM/IPC.h can be compiled and exercised without Eigen / Kokkos / Cabana.  It mimics
// only the interfaces the shim touches (written from the reference's documented layouts):
//   VECTOR<T,dim>      : T data[4], 16 bytes for int, 32 bytes for double   (Math/VECTOR.h:33-47)
//   MESH_NODE          : size, Get_Unchecked(i) -> tuple<VECTOR<T,3>&>, elements contiguous (32-byte stride)
//   MESH_NODE_ATTR     : Cabana-like AoSoA, bins of 4 nodes: x0[4][4] v[4][4] g[4][4] m[4]  (NOT constant stride)
//   FIELDS<...>::x0/v/g/m, Eigen::Triplet<T>, TIMER_FLAG
// The six contact templates below are what the shim renames to *_CPU; here they abort, which proves the
// <double,3,false,false> instantiation never reaches them.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <tuple>
#include <vector>

namespace Eigen {
template <class T> struct Triplet {
    int m_row, m_col; T m_value;
    Triplet() : m_row(0), m_col(0), m_value(0) {}
    Triplet(int r, int c, T v) : m_row(r), m_col(c), m_value(v) {}
    int row() const { return m_row; } int col() const { return m_col; } T value() const { return m_value; }
};
}
#define TIMER_FLAG(name) do { } while (0)

namespace JGSL {
template <class T, int dim> struct alignas(sizeof(T) * 4) VECTOR {
    T data[4];
    VECTOR() : data{0, 0, 0, 0} {}
    T& operator[](int d) { return data[d]; }
    const T& operator[](int d) const { return data[d]; }
};
template <class S> struct FIELDS;
template <class T, int dim> struct MESH_NODE {
    std::vector<VECTOR<T, dim>> v;
    size_t size = 0;
    std::tuple<VECTOR<T, dim>&> Get_Unchecked(size_t i) { return std::tuple<VECTOR<T, dim>&>(v[i]); }
};
template <class T, int dim> struct MESH_NODE_ATTR {
    struct Bin { VECTOR<T, dim> x0[4], vel[4], g[4]; T m[4]; };
    std::vector<Bin> bins;
    size_t size = 0;
    std::tuple<VECTOR<T, dim>&, VECTOR<T, dim>&, VECTOR<T, dim>&, T&> Get_Unchecked(size_t i)
    {
        Bin& b = bins[i / 4];
        return std::tuple<VECTOR<T, dim>&, VECTOR<T, dim>&, VECTOR<T, dim>&, T&>(b.x0[i % 4], b.vel[i % 4], b.g[i % 4], b.m[i % 4]);
    }
};
template <class T, int dim> struct FIELDS<MESH_NODE_ATTR<T, dim>> { enum INDICES { x0 = 0, v, g, m }; };

#define CIPC_STUB_ABORT(name) do { printf("reference CPU template %s reached in the shim harness\n", name); exit(3); } while (0)
template <class T, int dim, bool shell = false, bool elasticIPC = false>
void Compute_Constraint_Set(MESH_NODE<T, dim>&, MESH_NODE_ATTR<T, dim>&, const std::vector<int>&, const std::vector<VECTOR<int, 2>>&,
    const std::vector<VECTOR<int, 3>>&, const std::vector<int>&, const std::vector<VECTOR<int, 2>>&, const std::map<int, std::set<int>>&,
    const std::vector<T>&, const std::vector<T>&, const std::vector<T>&, const VECTOR<int, 2>&, const std::vector<bool>&, T, T, bool,
    std::vector<VECTOR<int, dim + 1>>&, std::vector<VECTOR<int, 2>>&, std::vector<VECTOR<T, 2>>&) { CIPC_STUB_ABORT("Compute_Constraint_Set"); }
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier(MESH_NODE<T, dim>&, MESH_NODE_ATTR<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&, const std::vector<VECTOR<T, 2>>&, T, T[], T, T&)
{ CIPC_STUB_ABORT("Compute_Barrier"); }
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier_Gradient(MESH_NODE<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&, const std::vector<VECTOR<T, 2>>&, T, T[], T, MESH_NODE_ATTR<T, dim>&)
{ CIPC_STUB_ABORT("Compute_Barrier_Gradient"); }
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier_Hessian(MESH_NODE<T, dim>&, MESH_NODE_ATTR<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&, const std::vector<VECTOR<T, 2>>&, T, T[], T,
    bool, std::vector<Eigen::Triplet<T>>&) { CIPC_STUB_ABORT("Compute_Barrier_Hessian"); }
template <class T, int dim, bool shell = false, bool elasticIPC = false>
void Compute_Intersection_Free_StepSize(MESH_NODE<T, dim>&, const std::vector<int>&, const std::vector<VECTOR<int, 2>>&, const std::vector<VECTOR<int, 3>>&,
    const std::vector<int>&, const std::vector<VECTOR<int, 2>>&, const std::map<int, std::set<int>>&, const VECTOR<int, 2>&, const std::vector<bool>&,
    const std::vector<T>&, T, T&) { CIPC_STUB_ABORT("Compute_Intersection_Free_StepSize"); }
template <class T, int dim, bool elasticIPC = false>
void Compute_Min_Dist2(MESH_NODE<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&, T, std::vector<T>&, T&) { CIPC_STUB_ABORT("Compute_Min_Dist2"); }
} // namespace JGSL
