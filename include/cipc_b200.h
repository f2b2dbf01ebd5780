/* cipc_b200.h -- C ABI of the B200-native C-IPC contact hot path (libcipc_b200.so).
 *
 * The reference (ipc-sim/Codim-IPC) has no FFI layer for this path: it is six header-only C++
 * function templates in Library/FEM/IPC.h instantiated inside the pybind11 module.  Each entry
 * point below is what a binding for one of those templates calls; `shim/FEM/IPC.h` in this repo
 * re-declares the six templates with the reference's signatures on top of this ABI (see
 * INTEGRATION.md).  All pointers are HOST pointers unless the name says `_dev`.  All functions
 * return a cipc_status; nothing falls back to the CPU -- without a CUDA device cipc_create fails.
 *
 * Reference interface replaced (file:line under /root/reference/Library):
 *   cipc_constraint_set   <- Compute_Constraint_Set              FEM/IPC.h:19-36   (3-D branch :143-661)
 *   cipc_barrier_energy   <- Compute_Barrier                     FEM/IPC.h:742-748
 *   cipc_barrier_gradient <- Compute_Barrier_Gradient            FEM/IPC.h:943-948
 *   cipc_barrier_hessian  <- Compute_Barrier_Hessian             FEM/IPC.h:1258-1265
 *   cipc_step_size        <- Compute_Intersection_Free_StepSize  FEM/IPC.h:1879-1890
 *   cipc_min_dist2        <- Compute_Min_Dist2                   FEM/IPC.h:2246-2249
 *   cipc_set_topology / cipc_set_positions / cipc_set_rest_positions / cipc_set_search_dir
 *                         <- the Storage->device marshalling of MESH_NODE / MESH_NODE_ATTR
 *                            (FEM/DATA_TYPE.h:7-31, Math/VECTOR.h:33-47: element i of a
 *                            VECTOR<double,3> store is 4 contiguous doubles = 32 bytes).
 */
#ifndef CIPC_B200_H
#define CIPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cipc_ctx cipc_ctx;

typedef enum {
    CIPC_OK = 0,
    CIPC_ERR_CUDA = 1,             /* CUDA runtime failure or no device; message in cipc_last_error */
    CIPC_ERR_NONPOSITIVE_DIST = 2, /* reference: printf("%le distance detected ...") + exit(-1)  IPC.h:773-776 */
    CIPC_ERR_ZERO_STEP = 3,        /* reference: coordinate dump + exit(-1)                   IPC.h:2014-2032 */
    CIPC_ERR_ARG = 4,
    CIPC_ERR_GRID = 5,             /* voxel grid does not fit the 30-bit cell key */
    CIPC_ERR_UNSUPPORTED = 6
} cipc_status;

/* 16-byte triplet, layout-identical to Eigen::Triplet<double,int> {int row; int col; double value;} */
typedef struct { int32_t row, col; double val; } cipc_triplet;

/* ---- lifetime ------------------------------------------------------------------------- */
/* device: CUDA ordinal.  rank/world: this process' share of the candidate-pair work (multi-GPU:
 * one process per GPU); rank=0, world=1 for a single GPU. */
int cipc_create(int device, int rank, int world, cipc_ctx** out);
/* One context over ndev GPUs of one box, driven by ONE calling thread (the reference calls the path from a single thread,
 * Shell/IMPLICIT_EULER.h:418-428): candidate pairs are partitioned by voxel slabs across the devices, helper threads overlap
 * the ranks' host round trips, exchanges go over NVLink peer copies / peer loads (gather + merge of the PP/PE stencils,
 * re-balancing of the constraint list into contiguous per-rank chunks, gradient sum); energy / step size / min distance are
 * combined on the host in rank order.  Supported on such a context: cipc_set_topology, the position / rest / search-direction
 * uploads, cipc_constraint_set, cipc_get/set_constraints[_strided], cipc_barrier_energy / _gradient / _hessian[_merged],
 * cipc_get_triplets, cipc_step_size, cipc_min_dist2, cipc_stage_ms (slowest rank), cipc_counter (sum), cipc_sync; every other
 * call returns CIPC_ERR_UNSUPPORTED.  Results equal the single-device context's (same constraint set as a sorted set; sums
 * in a different but fixed order). */
int cipc_create_multi(int ndev, const int* devices, cipc_ctx** out);
void cipc_destroy(cipc_ctx* ctx);
const char* cipc_last_error(cipc_ctx* ctx);

/* ---- marshalling ---------------------------------------------------------------------- */
/* Boundary primitive lists as the reference builds them every step (FEM/Shell/IMPLICIT_EULER.h:224-300).
 * be_stride / bt_stride: ints between consecutive edges / triangles (2 or 4; 3 or 4 -- VECTOR<int,k>
 * is 16 bytes, i.e. stride 4).  dbc: nV bytes (DBCb).  nnx: NNExclusion flattened to (key,member)
 * pairs.  Areas may be NULL unless elasticIPC is used.  Re-uploads only when the content changed. */
int cipc_set_topology(cipc_ctx* ctx, int nV, int nBN, const int32_t* BN, int nBE, const int32_t* BE, int be_stride,
                      int nBT, const int32_t* BT, int bt_stride, int nRod, const int32_t codimBNStartInd[2],
                      const uint8_t* dbc, int nNnxPairs, const int32_t* nnxPairs,
                      const double* BNArea, const double* BEArea, const double* BTArea);
/* stride_bytes: 32 for the reference's VECTOR<double,3> storage, 24 for packed xyz.  The four uploads below keep a 64-bit
 * content tag of what is resident: an array whose bytes equal the previous upload's is not sent again (the contact stage passes
 * the same X to up to eight consecutive calls). */
int cipc_set_positions(cipc_ctx* ctx, const double* X, int stride_bytes);
int cipc_set_rest_positions(cipc_ctx* ctx, const double* X0, int stride_bytes);
int cipc_set_search_dir(cipc_ctx* ctx, const double* p /* 3*nV packed, like std::vector<T> searchDir */);

/* ---- Compute_Constraint_Set ------------------------------------------------------------ */
/* Builds the constraint set for the resident positions; it stays resident on the device.
 * nC_out: number of constraints held by THIS rank. */
int cipc_constraint_set(cipc_ctx* ctx, int elasticIPC, double dHat2, double thickness, int* nC_out);
/* copies the resident set out: cs = nC x 4 int32 (VECTOR<int,4> stride 4), info = nC x 2 double */
int cipc_get_constraints(cipc_ctx* ctx, int32_t* cs, double* info);
/* makes a caller-owned set resident (used when the caller's vectors are not the last ones produced) */
int cipc_set_constraints(cipc_ctx* ctx, const int32_t* cs, const double* info, int nC);
/* the same two calls for the reference's std::vector<VECTOR<T,2>> stencilInfo, whose records are 32 bytes apart (VECTOR<T,dim>
 * always stores T data[4], Math/VECTOR.h:38-42): info_stride_bytes between consecutive (weight, dHat2) pairs */
int cipc_get_constraints_strided(cipc_ctx* ctx, int32_t* cs, double* info, int info_stride_bytes);
int cipc_set_constraints_strided(cipc_ctx* ctx, const int32_t* cs, const double* info, int info_stride_bytes, int nC);

/* ---- barrier terms on the resident constraint set and positions --------------------------- */
/* E_inout += sum_c w_c m_c b(d_c) e_c          (accumulates like the reference, IPC.h:940) */
int cipc_barrier_energy(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness, double* E_inout);
/* g[v] += ...  g_stride_bytes between consecutive nodes' 3 doubles (nodeAttr.g inside the AoSoA: see INTEGRATION.md) */
int cipc_barrier_gradient(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness,
                          double* g, int g_stride_bytes);
/* computes all (PSD-projected) blocks on the device; nTriplets_out = 144/81/36 per constraint */
int cipc_barrier_hessian(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness,
                         int projectSPD, int64_t* nTriplets_out);
/* delivers the triplets of the last cipc_barrier_hessian to host memory `out` (caller appends them, IPC.h:1371-1388).
 * Projected blocks travel over PCIe as compact factors (288 B instead of 2304 B per PT/EE stencil) and are expanded by
 * the host cores (CIPC_HOST_THREADS, default all); CIPC_TRIPLETS_DMA=1 copies the expanded device stream instead. */
int cipc_get_triplets(cipc_ctx* ctx, cipc_triplet* out);
/* Compute_Barrier_Hessian delivered as MERGED triplets.  The only consumer of the reference's triplet vector is
 * sysMtr.Construct_From_Triplet = Eigen setFromTriplets (Shell/INC_POTENTIAL.h:382, Math/CSR_MATRIX.h:49-56), which sums
 * entries with equal (row, col).  This call performs that sum on the device at 3x3-block granularity (upper block triangle,
 * mirrored on delivery): nTriplets_out = number of DISTINCT (row, col) entries of the contact matrix (119M instead of 909M
 * at 1M triangles), and the cipc_get_triplets that follows ships 80 bytes per unique upper block over PCIe.  The matrix
 * assembled from the merged triplets equals the one assembled from cipc_barrier_hessian's to summation order (1e-9 gate in
 * tests/test_gpu_merged.py).  Layout of the delivered array: the upper blocks in (row, col) order, then the mirrored blocks. */
int cipc_barrier_hessian_merged(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness,
                                int projectSPD, int64_t* nTriplets_out);
/* the same triplet stream resident in HBM (expanded on the device on first use); NULL on error */
cipc_triplet* cipc_dev_triplets(cipc_ctx* ctx);

/* ---- Compute_Intersection_Free_StepSize --------------------------------------------------- */
/* stepSize_inout: in = current upper bound, out = min(bound [possibly shrunk by the span rule,
 * Grid/SPATIAL_HASH.h:466-482], min over candidate pairs of the ACCD time of impact) of THIS rank's pairs */
int cipc_step_size(cipc_ctx* ctx, int elasticIPC, double thickness, double* stepSize_inout);

/* ---- Compute_Min_Dist2 ---------------------------------------------------------------------- */
/* dist2 may be NULL; minDist2 = min_c dist2[c] - thickness^2.  An empty resident set is a no-op (the reference returns
 * early and leaves its outputs untouched, IPC.h:2253). */
int cipc_min_dist2(cipc_ctx* ctx, double thickness, double* dist2, double* minDist2);

/* ---- lagged friction: FEM/FRICTION.h, FEM/FRICTION_UTILS.h ------------------------------- */
/* previous-step positions Xn (MESH_NODE storage, stride 32; 24 for packed xyz) */
int cipc_set_prev_positions(cipc_ctx* ctx, const double* Xn, int stride_bytes);
/* Compute_Friction_Basis<T,3,elasticIPC> (FRICTION.h:16-124) on the RESIDENT contact constraint set and positions:
 * drops the mollified stencils (order of the others preserved), computes closestPoint, tanBasis and the lagged
 * normalForce = -b'(d - xi^2) * 2 sqrt(d) * w.  The friction set stays resident; nF_out = its size. */
int cipc_friction_basis(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness, int* nF_out);
/* copies the resident friction set out in the reference's container layouts: fcs nF x 4 int32 (VECTOR<int,4>),
 * closestPoint nF x 2 double (Eigen::Matrix<T,2,1>), tanBasis nF x 6 double (Eigen::Matrix<T,3,2>, column-major),
 * normalForce nF double.  Any pointer may be NULL. */
int cipc_get_friction_basis(cipc_ctx* ctx, int32_t* fcs, double* closestPoint, double* tanBasis, double* normalForce);
/* makes a caller-owned friction set resident (when the caller's vectors are not the last ones produced);
 * closestPoint / tanBasis may be NULL when only cipc_friction_coef follows */
int cipc_set_friction_basis(cipc_ctx* ctx, const int32_t* fcs, const double* closestPoint, const double* tanBasis,
                            const double* normalForce, int nF);
/* Compute_Friction_Coef (FRICTION.h:126-170): resident normalForce[c] *= muComp[comp(v0) + comp(v1) * nComp]; *mu_out = 1 */
int cipc_friction_coef(cipc_ctx* ctx, int nComp, const int32_t* compNodeRange, const double* muComp /* nComp x nComp */, double* mu_out);
/* Compute_Friction_Potential (FRICTION.h:172-252): E_inout += mu * sum_c mult_c normalForce_c f0(|u_c|) */
int cipc_friction_energy(cipc_ctx* ctx, double epsvh2, double mu, double* E_inout);
/* Compute_Friction_Gradient (FRICTION.h:254-379): g[v] += ... (same stride convention as cipc_barrier_gradient) */
int cipc_friction_gradient(cipc_ctx* ctx, double epsvh2, double mu, double* g, int g_stride_bytes);
/* Compute_Friction_Hessian (FRICTION.h:381-663): 144/81/36 triplets per friction stencil; deliver them with
 * cipc_get_triplets / cipc_dev_triplets exactly like the barrier Hessian's (the caller appends) */
int cipc_friction_hessian(cipc_ctx* ctx, double epsvh2, double mu, int projectSPD, int64_t* nTriplets_out);
/* Compute_Friction_Hessian delivered as merged triplets (see cipc_barrier_hessian_merged) */
int cipc_friction_hessian_merged(cipc_ctx* ctx, double epsvh2, double mu, int projectSPD, int64_t* nTriplets_out);
/* device-resident variants: energy -> cipc_dev_scalars()[4]; gradient -> cipc_dev_gradient(), added to what is there
 * when accumulate != 0 (e.g. after cipc_barrier_gradient_dev) */
int cipc_friction_energy_dev(cipc_ctx* ctx, double epsvh2, double mu);
int cipc_friction_gradient_dev(cipc_ctx* ctx, double epsvh2, double mu, int accumulate);
/* Compute_Friction_Hessian with the triplet stream left in HBM (cipc_dev_triplets), fused like cipc_barrier_hessian_dev */
int cipc_friction_hessian_dev(cipc_ctx* ctx, double epsvh2, double mu, int projectSPD, int64_t* nTriplets_out);

/* ---- device-resident line search ------------------------------------------------------------
 * Shell/IMPLICIT_EULER.h:102-131 evaluates X = Xprev + alpha p, Compute_Constraint_Set and Compute_Min_Dist2 per
 * trial step; with these three calls the trial positions never cross PCIe:
 *   cipc_save_positions  : Xprev <- the resident positions (call once per line search, after cipc_set_positions)
 *   cipc_step_positions  : resident X <- Xprev + alpha * (resident search direction), unfused multiply-add
 *   cipc_get_positions   : copies the resident positions out (stride 32 or 24) once a step has been accepted */
int cipc_save_positions(cipc_ctx* ctx);
int cipc_step_positions(cipc_ctx* ctx, double alpha);
int cipc_get_positions(cipc_ctx* ctx, double* X, int stride_bytes);

/* ---- Hessian triplets -> CSR on the device ---------------------------------------------------
 * What Math/CSR_MATRIX.h:49-56 Construct_From_Triplet (Eigen setFromTriplets; Shell/INC_POTENTIAL.h:382-394) does with
 * the contact / friction triplets, done in HBM at 3x3-block granularity: duplicates summed (in a reproducible order),
 * column indices sorted inside every row, explicit zeros kept.  The matrix is 3nV x 3nV.
 *   cipc_csr_begin  : start a new matrix
 *   cipc_csr_add    : append the blocks of the Hessian computed last (cipc_barrier_hessian[_dev] / cipc_friction_hessian);
 *                     call it after each Hessian that should be part of the matrix (the triplet stream is reused)
 *   cipc_csr_finish : sort, merge and emit; nnz_out = stored entries
 *   cipc_get_csr    : rowPtr (3nV+1), colIdx (nnz), val (nnz) -> host (Eigen outerIndexPtr / innerIndexPtr / valuePtr of
 *                     a row-major SparseMatrix<double,RowMajor,int>); any pointer may be NULL */
int cipc_csr_begin(cipc_ctx* ctx);
int cipc_csr_add(cipc_ctx* ctx);
int cipc_csr_finish(cipc_ctx* ctx, int64_t* nnz_out);
int cipc_get_csr(cipc_ctx* ctx, int32_t* rowPtr, int32_t* colIdx, double* val);

/* ---- boundary-primitive construction --------------------------------------------------------
 * Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834: a std::map over the 3T directed triangle edges, rebuilt
 * every time step by Shell/IMPLICIT_EULER.h:224-243) plus the seg / rod / particle appends of Shell/IMPLICIT_EULER.h:245-277,
 * on the device, with the reference's exact output order:
 *   boundaryTri = tri in order; boundaryEdge = undirected edges oriented like their first mention, in lexicographic order of
 *   the oriented pair, then seg, then rod; boundaryNode = ascending surface vertices of non-zero area, then both ends of
 *   every seg, [codim0] rod nodes ascending, [codim1] particles; BNArea (surface + rod nodes only: nBNArea values), BEArea
 *   (surface + rod edges only: nBE - nSeg values, like the reference's vector), BTArea.
 * X: nV node positions (stride 32 or 24 bytes); tri / seg / rod: int records with the given stride in ints (3|4, 2|4, 2|4);
 * rodRadius[i] = rodInfo[i][2].  counts_out = {nBN, nBE, nBT, codimBNStartInd[0], codimBNStartInd[1], nBNArea}.
 * The result is cached on a content hash of every input: an unchanged call costs the hash only.
 * cipc_get_boundary copies the lists out (any pointer may be NULL; be_stride / bt_stride as in cipc_set_topology). */
int cipc_build_boundary(cipc_ctx* ctx, int nV, const double* X, int x_stride_bytes, int nTri, const int32_t* tri, int tri_stride,
                        int nSeg, const int32_t* seg, int seg_stride, int nRod, const int32_t* rod, int rod_stride,
                        const double* rodRadius, int nParticle, const int32_t* particle, int32_t counts_out[6]);
int cipc_get_boundary(cipc_ctx* ctx, int32_t* BN, int32_t* BE, int be_stride, int32_t* BT, int bt_stride, double* BNArea,
                      double* BEArea, double* BTArea);

/* ---- device-resident access (multi-GPU reductions, benchmarking) ----------------------------- */
double* cipc_dev_positions(cipc_ctx* ctx);      /* nV x 4 doubles (x,y,z,pad) */
double* cipc_dev_gradient(cipc_ctx* ctx);       /* 3*nV doubles written by the last cipc_barrier_gradient[_hessian]_dev; after
                                                 * cipc_barrier_gradient_hessian_dev slot [3*nV] holds the energy of the last
                                                 * cipc_barrier_energy_dev, so one all-reduce(sum) of 3*nV+1 doubles covers both */
double* cipc_dev_scalars(cipc_ctx* ctx);        /* [0]=barrier energy, [1]=step size, [2]=min dist2, [4]=friction energy of the last *_dev call */
/* same stages with every result left on the device (no D2H): */
int cipc_barrier_energy_dev(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness);
int cipc_barrier_gradient_dev(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness);
/* Compute_Barrier_Hessian with the (row, col, value) stream left in HBM (cipc_dev_triplets): factoring and expansion
 * run fused in one kernel per stencil class, the factors never travel through HBM.  cipc_get_triplets afterwards copies
 * the expanded stream over PCIe. */
int cipc_barrier_hessian_dev(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness,
                             int projectSPD, int64_t* nTriplets_out);
/* Compute_Barrier_Gradient + Compute_Barrier_Hessian (projectSPD = true) in one pass over the stencils, as the Newton
 * iteration evaluates them back to back (Shell/IMPLICIT_EULER.h:464, INC_POTENTIAL.h:374): the gradient (cipc_dev_gradient,
 * overwritten) is accumulated by the fused Hessian kernels from the positions they already hold */
int cipc_barrier_gradient_hessian_dev(cipc_ctx* ctx, int elasticIPC, double dHat2, const double kappa[3], double thickness,
                                      int64_t* nTriplets_out);
int cipc_step_size_dev(cipc_ctx* ctx, int elasticIPC, double thickness, double stepSize_in);
int cipc_min_dist2_dev(cipc_ctx* ctx, double thickness);
int cipc_sync(cipc_ctx* ctx);
/* run all subsequent work on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream) so that
 * the caller's events / NCCL collectives order against it without host synchronisation; NULL restores the
 * library's own stream */
int cipc_set_stream(cipc_ctx* ctx, void* cuda_stream);
/* stage timers (cipc_stage_ms) record two CUDA events per stage scope; on = 0 switches them off (cipc_stage_ms then returns -1),
 * on != 0 (default) back on */
int cipc_set_timing(cipc_ctx* ctx, int on);
/* CUDA events on the library's stream: record into slot [0,64) / elapsed milliseconds between two slots */
int cipc_event_record(cipc_ctx* ctx, int slot);
double cipc_event_elapsed_ms(cipc_ctx* ctx, int slot_a, int slot_b);

/* ---- introspection ------------------------------------------------------------------------ */
/* device milliseconds (CUDA events on the library's stream) of the named stage of the last call:
 * "ccs_hash_build","ccs_pairs","ccs_narrow","ccs_merge","barrier_E","barrier_g","barrier_H",
 * "ccd_hash_build","ccd_pairs","ccd_accd","min_dist"; per-kind sub-scopes "ccs_narrow_pt","ccs_narrow_ee","ccd_accd_pt",
 * "ccd_accd_ee" ; -1 if unknown */
double cipc_stage_ms(cipc_ctx* ctx, const char* stage);
/* counters of the last call: "candidates_pt","candidates_ee","candidates_pe","candidates_pp",
 * "hash_entries","hash_cells","constraints","ccd_pairs" ; -1 if unknown */
int64_t cipc_counter(cipc_ctx* ctx, const char* name);
int64_t cipc_kernel_launches(void);  /* kernels launched by this library since load */
/* 64-bit content hash of a host array, computed by a few host threads (~50 GB/s): the shim's change detector for the
 * caller-owned containers (constraint set, friction set) it keeps resident on the device between calls */
uint64_t cipc_hash_bytes(const void* p, size_t n);
/* starts touching the pages of [p, p + bytes) on the host thread pool in the background and returns at once: a caller that
 * knows roughly how many triplets the next Hessian delivers (the previous Newton iteration's count) reserves its vector first
 * and lets the first-touch page faults of the fresh allocation overlap the device work; cipc_get_triplets waits for it. */
int cipc_host_prefault_async(void* p, size_t bytes);
/* runs fn(begin, end, user) over [0, n) in pieces of `grain` on the library's persistent host thread pool (CIPC_HOST_THREADS,
 * default all cores) and returns when all pieces are done: the shim's loops over the reference's AoSoA node storage
 * (gradient accumulation nodeAttr.g +=, rest-position gather) use it.  Not re-entrant from inside fn. */
typedef void (*cipc_range_fn)(size_t begin, size_t end, void* user);
int cipc_host_parallel_for(size_t n, size_t grain, cipc_range_fn fn, void* user);
const char* cipc_version(void);

/* ---- test hooks (device primitives, exercised by tests/) ----------------------------------- */
int cipc_test_scan(cipc_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total);
int cipc_test_sort(cipc_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int bits);
int cipc_test_make_pd(cipc_ctx* ctx, double* H /* count x n x n */, int n, int count);

#ifdef __cplusplus
}
#endif
#endif /* CIPC_B200_H */
