// TEST INFRASTRUCTURE (oracle/_ref build only): shadows the reference's Utils/MESHIO.h (OBJ/VTK file
// I/O, mesh subdivision, pybind registration -- 1.4K lines that need Boost and the full pybind11) with
// just the includes through which FEM/IPC.h and Grid/SPATIAL_HASH.h receive their types.  Nothing of
// MESHIO.h is on the contact path.
#pragma once
#include <Math/VECTOR.h>
#include <set>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <deque>
#include <numeric>
#include <sstream>
#include <fstream>
#include <FEM/DATA_TYPE.h>
#include <Grid/SPATIAL_HASH.h>
