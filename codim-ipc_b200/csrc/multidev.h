// multidev.h -- several GPUs of one box behind ONE context and ONE calling thread (cipc_create_multi): what a C++ caller of the
// reference's six templates links against to use the whole node (the reference calls the path from a single thread,
// Shell/IMPLICIT_EULER.h:418-428, INC_POTENTIAL.h:373-382).
//
// Every device gets an ordinary single-device context with (rank, world) = (r, N): the candidate pairs are partitioned by
// voxel slabs exactly as in the one-process-per-GPU mode.  Each rank has a helper thread that issues
// its launches and waits on its stream, so the ranks' host round trips overlap; the caller's thread drives rank 0.
// Exchanges stay on NVLink (peer copies / peer loads, no host staging):
//   constraint set : PT / EE / mollified stencils are unique to the rank that enumerated the pair.  The PP / PE stencils every
//                    rank de-duplicated locally are gathered on device 0 and merged with their multiplicities (IPC.h:599-654
//                    keys on the raw tuple, the same key can come from pairs of different slabs); a merged stencil is owned by
//                    the rank whose record claimed it.  The global list G = [rank 0's | rank 0's merged PP/PE | rank 1's | ...]
//                    is then re-cut into N equal CONTIGUOUS chunks, one per rank (peer copies), so the barrier terms are
//                    balanced, every rank's outputs (constraints, dist2, triplets) are one contiguous slice of the caller's
//                    containers, and the chunks stay aligned with the voxel slabs (few Hessian blocks shared between ranks).
//   energy / step / min distance : one scalar per rank, combined on the host in rank order.
//   gradient       : rank 0 sums the ranks' 3 nV vectors with peer loads (k_sum_peers), one D2H.
//   Hessian        : every rank delivers the (merged) triplets of its chunk into its slice of the caller's vector; entries
//                    with equal (row, col) from different ranks are summed by the consumer like any duplicate triplet.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct cipc_multi;
namespace cipc {

constexpr int MAX_RANKS = 16;
struct PeerPtrs { const double* p[MAX_RANKS]; int n; };
__global__ void k_sum_peers(double* __restrict__ dst, PeerPtrs src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double s = dst[i];
        for (int r = 0; r < src.n; ++r) s += src.p[r][i]; // rank order: reproducible
        dst[i] = s;
    }
}
// weighted variant of k_dedup_insert: records carry their multiplicity in -c.w (locally de-duplicated PP / PE stencils)
__global__ void k_dedup_insert_w(const int4* __restrict__ raw, u32 n, u32* slots, u32* slotCnt, u32 mask)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 k = raw[i];
    const u32 w = (u32)(-k.w);
    u32 h = hash3(k.x, k.y, k.z) & mask;
    while (true) {
        const u32 prev = atomicCAS(&slots[h], 0xffffffffu, i);
        if (prev == 0xffffffffu) { atomicAdd(&slotCnt[h], w); return; }
        const int4 o = raw[prev];
        if (o.x == k.x && o.y == k.y && o.z == k.z) { atomicAdd(&slotCnt[h], w); return; }
        h = (h + 1) & mask;
    }
}

// Emission of the merged PP / PE stencils grouped by OWNER rank = the rank whose record claimed the hash slot (rankOff: offsets
// of the ranks' records in the gathered list).  PASS 0 counts per owner, PASS 1 writes at ownerBase[o] + running counter: the
// global list then reads [rank 0's stencils | rank 0's merged PP/PE | rank 1's ... ], so the re-cut chunks stay aligned with the
// ranks' voxel slabs and few 3x3 Hessian blocks are shared between ranks.
struct RankOffsets { u32 off[MAX_RANKS + 1]; int n; };
template <int PASS>
__global__ void k_dedup_emit_owned(const int4* __restrict__ raw, const u32* __restrict__ slots, const u32* __restrict__ slotCnt, u32 nSlots,
    RankOffsets ro, const u32* __restrict__ ownerBase, u32* __restrict__ ownerCnt, int4* __restrict__ out)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nSlots) return;
    const u32 r = slots[s];
    if (r == 0xffffffffu) return;
    int o = 0;
    while (o + 1 < ro.n && r >= ro.off[o + 1]) ++o;
    const u32 k = atomicAdd(&ownerCnt[o], 1u);
    if (PASS == 1) {
        const int4 q = raw[r];
        out[ownerBase[o] + k] = make_int4(q.x, q.y, q.z, -(int)slotCnt[s]);
    }
}

// one helper thread per rank >= 1: runs the jobs the calling thread hands it
class RankThread {
public:
    RankThread() : th_([this] { loop(); }) {}
    ~RankThread()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        th_.join();
    }
    void start(std::function<void()> f)
    {
        { std::lock_guard<std::mutex> lk(m_); job_ = std::move(f); busy_ = true; }
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return !busy_; });
    }
private:
    void loop()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || busy_; });
                if (stop_) return;
                f = job_;
            }
            f();
            { std::lock_guard<std::mutex> lk(m_); busy_ = false; }
            done_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void()> job_;
    bool busy_ = false, stop_ = false;
    std::thread th_;
};

} // namespace cipc

struct cipc_multi {
    int n = 0;
    std::vector<cipc_ctx*> sub;                         // sub[r]: device devs[r], rank r of n
    std::vector<std::unique_ptr<cipc::RankThread>> thr; // thr[r - 1] serves rank r
    std::vector<size_t> chunk;                          // n + 1 offsets of the ranks' chunks in the global constraint list
    std::vector<int64_t> tripCount;                     // triplets of the last Hessian per rank
    cipc::DevBuf<int4> mergeTmp;                        // device 0: merged PP / PE stencils
    cipc::DevBuf<cipc::u32> ownerDev;                   // device 0: per-owner counters [0, MAX_RANKS] and bases [MAX_RANKS + 1, ...]
    // runs f(r) for every rank (rank 0 on the calling thread) and returns the first non-OK status
    int run_all(const std::function<int(int)>& f)
    {
        std::vector<int> st(n, 0);
        for (int r = 1; r < n; ++r) thr[r - 1]->start([&, r] { st[r] = f(r); });
        st[0] = f(0);
        for (int r = 1; r < n; ++r) thr[r - 1]->wait();
        for (int r = 0; r < n; ++r) if (st[r]) return st[r];
        return 0;
    }
};
