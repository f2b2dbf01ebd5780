// TEST INFRASTRUCTURE: drives codim-ipc_b200/shim/FEM/FRICTION.h in the reference's call order
// (Library/FEM/Shell/IMPLICIT_EULER.h:419-464, INC_POTENTIAL.h:374-376): constraint set -> friction basis -> coef ->
// potential -> gradient -> Hessian, with the reference's container types; results are dumped for the Python test.
//   usage: friction_harness <scene.bin> <out.bin>     (scene.bin of test_gpu_shim.py followed by Xn)
#include <FEM/FRICTION.h> // resolves to the shim
#include <cstdint>
#include <cstdio>
#include <fstream>

using namespace JGSL;
template <class T> static void rd(std::ifstream& f, T* p, size_t n) { f.read((char*)p, sizeof(T) * n); }
template <class T> static void wr(std::ofstream& f, const T* p, size_t n) { f.write((const char*)p, sizeof(T) * n); }

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    int32_t hdr[8]; // nV nBN nBE nBT nRod codim0 codim1 nNnx
    rd(f, hdr, 8);
    const int nV = hdr[0], nBN = hdr[1], nBE = hdr[2], nBT = hdr[3], nRod = hdr[4], nNnx = hdr[7];
    double par[6]; // dHat2 xi kappa0 kappa1 kappa2 stepSize
    rd(f, par, 6);
    std::vector<double> Xf(3 * nV), X0f(3 * nV), pf(3 * nV), Xnf(3 * nV);
    rd(f, Xf.data(), 3 * nV); rd(f, X0f.data(), 3 * nV); rd(f, pf.data(), 3 * nV);
    std::vector<int32_t> BN(nBN), BE(2 * nBE), BT(3 * nBT), nnx(2 * nNnx);
    std::vector<uint8_t> dbc(nV);
    rd(f, BN.data(), nBN); rd(f, BE.data(), 2 * nBE); rd(f, BT.data(), 3 * nBT); rd(f, dbc.data(), nV); rd(f, nnx.data(), 2 * nNnx);
    rd(f, Xnf.data(), 3 * nV);

    MESH_NODE<double, 3> X, Xn;
    MESH_NODE_ATTR<double, 3> nodeAttr;
    X.v.resize(nV); X.size = nV; Xn.v.resize(nV); Xn.size = nV;
    nodeAttr.bins.resize((nV + 3) / 4); nodeAttr.size = nV;
    for (int i = 0; i < nV; ++i)
        for (int d = 0; d < 3; ++d) {
            X.v[i][d] = Xf[3 * i + d]; Xn.v[i][d] = Xnf[3 * i + d];
            std::get<0>(nodeAttr.Get_Unchecked(i))[d] = X0f[3 * i + d];
            std::get<2>(nodeAttr.Get_Unchecked(i))[d] = 0.5;
        }
    std::vector<int> boundaryNode(BN.begin(), BN.end()), particle;
    std::vector<VECTOR<int, 2>> boundaryEdge(nBE), rod(nRod);
    std::vector<VECTOR<int, 3>> boundaryTri(nBT);
    for (int i = 0; i < nBE; ++i) { boundaryEdge[i][0] = BE[2 * i]; boundaryEdge[i][1] = BE[2 * i + 1]; }
    for (int i = 0; i < nRod; ++i) rod[i] = boundaryEdge[nBE - nRod + i];
    for (int i = 0; i < nBT; ++i) for (int d = 0; d < 3; ++d) boundaryTri[i][d] = BT[3 * i + d];
    for (int i = hdr[6]; i < nBN; ++i) particle.push_back(BN[i]);
    std::map<int, std::set<int>> NNExclusion;
    for (int i = 0; i < nNnx; ++i) NNExclusion[nnx[2 * i]].insert(nnx[2 * i + 1]);
    VECTOR<int, 2> codim; codim[0] = hdr[5]; codim[1] = hdr[6];
    std::vector<bool> DBCb(nV);
    for (int i = 0; i < nV; ++i) DBCb[i] = dbc[i] != 0;
    std::vector<double> BNArea, BEArea, BTArea;
    double dHat2 = par[0], xi = par[1], kappa[3] = {par[2], par[3], par[4]};

    std::vector<VECTOR<int, 4>> constraintSet, fricConstraintSet(3);
    std::vector<VECTOR<int, 2>> cs_PTEE;
    std::vector<VECTOR<double, 2>> stencilInfo;
    Compute_Constraint_Set<double, 3, false, false>(X, nodeAttr, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion, BNArea, BEArea,
        BTArea, codim, DBCb, dHat2, xi, false, constraintSet, cs_PTEE, stencilInfo);
    std::vector<Eigen::Matrix<double, 2, 1>> closestPoint;
    std::vector<Eigen::Matrix<double, 3, 2>> tanBasis;
    std::vector<double> normalForce;
    Compute_Friction_Basis<double, 3, false>(X, constraintSet, stencilInfo, fricConstraintSet, closestPoint, tanBasis, normalForce, dHat2, kappa, xi);
    std::vector<double> nf0 = normalForce;
    std::vector<int> compNodeRange = {nV / 2, nV};
    std::vector<double> muComp = {0.3, 0.7, 0.5, 0.2};
    double mu = 0.0;
    Compute_Friction_Coef<double, 3>(fricConstraintSet, compNodeRange, muComp, normalForce, mu);
    const double epsvh2 = 1e-10;
    double E = 2.0; // accumulates
    Compute_Friction_Potential(X, Xn, fricConstraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, E);
    Compute_Friction_Gradient(X, Xn, fricConstraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, nodeAttr);
    std::vector<Eigen::Triplet<double>> triplets(5, Eigen::Triplet<double>(1, 2, 3.0)); // appends
    Compute_Friction_Hessian(X, Xn, fricConstraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, true, triplets);
    // caller-made copies (different vectors, forces the explicit upload path), reversed order
    std::vector<VECTOR<int, 4>> csCopy(fricConstraintSet.rbegin(), fricConstraintSet.rend());
    std::vector<Eigen::Matrix<double, 2, 1>> cpCopy(closestPoint.rbegin(), closestPoint.rend());
    std::vector<Eigen::Matrix<double, 3, 2>> tbCopy(tanBasis.rbegin(), tanBasis.rend());
    std::vector<double> nfCopy(normalForce.rbegin(), normalForce.rend());
    double E2 = 0;
    Compute_Friction_Potential(X, Xn, csCopy, cpCopy, tbCopy, nfCopy, epsvh2, mu, E2);

    std::ofstream o(argv[2], std::ios::binary);
    int64_t n = (int64_t)constraintSet.size(), nF = (int64_t)fricConstraintSet.size(), nt = (int64_t)triplets.size();
    wr(o, &n, 1); wr(o, &nF, 1); wr(o, &nt, 1);
    for (auto& c : constraintSet) wr(o, c.data, 4);
    for (auto& s : stencilInfo) wr(o, s.data, 2);
    for (auto& c : fricConstraintSet) wr(o, c.data, 4);
    for (auto& c : closestPoint) wr(o, c.data(), 2);
    for (auto& c : tanBasis) wr(o, c.data(), 6);
    wr(o, nf0.data(), nf0.size());
    wr(o, normalForce.data(), normalForce.size());
    double sc[3] = {E, E2, mu};
    wr(o, sc, 3);
    for (int i = 0; i < nV; ++i) wr(o, std::get<2>(nodeAttr.Get_Unchecked(i)).data, 3);
    for (auto& t : triplets) { int32_t rc[2] = {t.row(), t.col()}; double v = t.value(); wr(o, rc, 2); wr(o, &v, 1); }
    printf("friction harness ok: %lld constraints, %lld friction stencils, %lld triplets, E=%.6e\n", (long long)n, (long long)nF, (long long)nt, E);
    return 0;
}
