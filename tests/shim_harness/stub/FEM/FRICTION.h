// TEST INFRASTRUCTURE: stand-in for the reference's Library/FEM/FRICTION.h (see stub/FEM/IPC.h): the five friction
// templates the shim renames to *_CPU abort here, which proves the <double,3> instantiation never reaches them.
// Eigen::Matrix<T,R,C> is mimicked as the packed column-major record Eigen uses for small fixed sizes.
#pragma once
#include <FEM/IPC.h>

namespace Eigen {
template <class T, int R, int C> struct Matrix {
    T m[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) m[i] = T(0); }
    T* data() { return m; }
    const T* data() const { return m; }
    T& operator[](int i) { return m[i]; }
    const T& operator[](int i) const { return m[i]; }
    T& operator()(int r, int c) { return m[r + R * c]; }
    const T& operator()(int r, int c) const { return m[r + R * c]; }
};
}

namespace JGSL {
template <class T, int dim = 3, bool elasticIPC = false>
void Compute_Friction_Basis(MESH_NODE<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&, const std::vector<VECTOR<T, 2>>&,
    std::vector<VECTOR<int, dim + 1>>&, std::vector<Eigen::Matrix<T, dim - 1, 1>>&, std::vector<Eigen::Matrix<T, dim, dim - 1>>&, std::vector<T>&, T, T[], T)
{ CIPC_STUB_ABORT("Compute_Friction_Basis"); }
template <class T, int dim = 3>
void Compute_Friction_Coef(const std::vector<VECTOR<int, dim + 1>>&, const std::vector<int>&, const std::vector<T>&, std::vector<T>&, T&)
{ CIPC_STUB_ABORT("Compute_Friction_Coef"); }
template <class T, int dim = 3>
void Compute_Friction_Potential(MESH_NODE<T, dim>&, MESH_NODE<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&,
    const std::vector<Eigen::Matrix<T, dim - 1, 1>>&, const std::vector<Eigen::Matrix<T, dim, dim - 1>>&, const std::vector<T>&, T, T, T&)
{ CIPC_STUB_ABORT("Compute_Friction_Potential"); }
template <class T, int dim = 3>
void Compute_Friction_Gradient(MESH_NODE<T, dim>&, MESH_NODE<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&,
    const std::vector<Eigen::Matrix<T, dim - 1, 1>>&, const std::vector<Eigen::Matrix<T, dim, dim - 1>>&, const std::vector<T>&, T, T,
    MESH_NODE_ATTR<T, dim>&)
{ CIPC_STUB_ABORT("Compute_Friction_Gradient"); }
template <class T, int dim = 3>
void Compute_Friction_Hessian(MESH_NODE<T, dim>&, MESH_NODE<T, dim>&, const std::vector<VECTOR<int, dim + 1>>&,
    const std::vector<Eigen::Matrix<T, dim - 1, 1>>&, const std::vector<Eigen::Matrix<T, dim, dim - 1>>&, const std::vector<T>&, T, T, bool,
    std::vector<Eigen::Triplet<T>>&)
{ CIPC_STUB_ABORT("Compute_Friction_Hessian"); }
}
