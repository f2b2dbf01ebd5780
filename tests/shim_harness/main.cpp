// TEST INFRASTRUCTURE: drives codim-ipc_b200/shim/FEM/IPC.h exactly like the reference's time stepper does
// (Library/FEM/Shell/IMPLICIT_EULER.h:97-131,331,419 and INC_POTENTIAL.h:148,273,374): the six templates are
// called with the reference's container types and the results are dumped for the Python test to compare with the
// oracle.   usage: shim_harness <scene.bin> <out.bin>
#include <FEM/IPC.h> // resolves to the shim; the shim include_next's the stub "reference" header
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
    std::vector<double> Xf(3 * nV), X0f(3 * nV), pf(3 * nV);
    rd(f, Xf.data(), 3 * nV); rd(f, X0f.data(), 3 * nV); rd(f, pf.data(), 3 * nV);
    std::vector<int32_t> BN(nBN), BE(2 * nBE), BT(3 * nBT), nnx(2 * nNnx);
    std::vector<uint8_t> dbc(nV);
    rd(f, BN.data(), nBN); rd(f, BE.data(), 2 * nBE); rd(f, BT.data(), 3 * nBT); rd(f, dbc.data(), nV); rd(f, nnx.data(), 2 * nNnx);

    MESH_NODE<double, 3> X;
    MESH_NODE_ATTR<double, 3> nodeAttr;
    X.v.resize(nV); X.size = nV;
    nodeAttr.bins.resize((nV + 3) / 4); nodeAttr.size = nV;
    for (int i = 0; i < nV; ++i)
        for (int d = 0; d < 3; ++d) {
            X.v[i][d] = Xf[3 * i + d];
            std::get<0>(nodeAttr.Get_Unchecked(i))[d] = X0f[3 * i + d];
            std::get<2>(nodeAttr.Get_Unchecked(i))[d] = 0.5; // pre-existing gradient content must be kept (+=)
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

    std::vector<VECTOR<int, 4>> constraintSet(7); // stale content must be discarded
    std::vector<VECTOR<int, 2>> cs_PTEE;
    std::vector<VECTOR<double, 2>> stencilInfo;
    Compute_Constraint_Set<double, 3, false, false>(X, nodeAttr, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion, BNArea, BEArea,
        BTArea, codim, DBCb, dHat2, xi, false, constraintSet, cs_PTEE, stencilInfo);
    double E = 2.0; // accumulates
    Compute_Barrier<double, 3, false>(X, nodeAttr, constraintSet, stencilInfo, dHat2, kappa, xi, E);
    Compute_Barrier_Gradient<double, 3, false>(X, constraintSet, stencilInfo, dHat2, kappa, xi, nodeAttr);
    std::vector<Eigen::Triplet<double>> triplets(5, Eigen::Triplet<double>(1, 2, 3.0)); // appends
    Compute_Barrier_Hessian<double, 3, false>(X, nodeAttr, constraintSet, stencilInfo, dHat2, kappa, xi, true, triplets);
    double stepSize = par[5];
    Compute_Intersection_Free_StepSize<double, 3, false, false>(X, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion, codim, DBCb, pf, xi,
        stepSize);
    std::vector<double> dist2;
    double minDist2 = 0;
    Compute_Min_Dist2<double, 3, false>(X, constraintSet, xi, dist2, minDist2);
    // a caller-made copy of the set (different vector) must also work: forces the explicit upload path
    std::vector<VECTOR<int, 4>> csCopy(constraintSet.rbegin(), constraintSet.rend());
    std::vector<VECTOR<double, 2>> infoCopy(stencilInfo.rbegin(), stencilInfo.rend());
    double E2 = 0;
    Compute_Barrier<double, 3, false>(X, nodeAttr, csCopy, infoCopy, dHat2, kappa, xi, E2);

    std::ofstream o(argv[2], std::ios::binary);
    int64_t n = (int64_t)constraintSet.size(), nt = (int64_t)triplets.size();
    wr(o, &n, 1); wr(o, &nt, 1);
    for (auto& c : constraintSet) wr(o, c.data, 4);
    for (auto& s : stencilInfo) wr(o, s.data, 2);
    double sc[4] = {E, stepSize, minDist2, E2};
    wr(o, sc, 4);
    for (int i = 0; i < nV; ++i) wr(o, std::get<2>(nodeAttr.Get_Unchecked(i)).data, 3);
    for (auto& t : triplets) { int32_t rc[2] = {t.row(), t.col()}; double v = t.value(); wr(o, rc, 2); wr(o, &v, 1); }
    wr(o, dist2.data(), dist2.size());
    printf("shim harness ok: %lld constraints, %lld triplets, E=%.6e step=%.6f\n", (long long)n, (long long)nt, E, stepSize);
    return 0;
}
