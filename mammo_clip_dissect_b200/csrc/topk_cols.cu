// K2 -- per-neuron top-k over the probe-image axis (replaces torch.topk(A, dim=0, k),
// reference concept_vit/similarity.py:55 / :82 / :107).
//
// A is [N images, K neurons] row-major, so a neuron's activations are strided by K floats.
// One consumer thread owns one neuron column; a CTA owns COLS adjacent columns and one slice of
// the image axis.  A producer warp streams [ROWS x COLS] tiles of A into a shared-memory ring
// with 1-D bulk async copies (TMA engine, mbarrier completion, L2 evict-first: A is read once),
// so the bytes in flight per SM do not depend on how many consumer warps there are.
//
// Per column the k best (value, index) pairs live in a shared-memory min-heap (root = current
// k-th best).  The scan compares every element with the root's value held in a register; the
// ~k*ln(N/k) elements that beat it are appended to a small per-column pending list, and when any
// lane's list is nearly full the warp folds its lists into the heaps in lock step.  Scanning in
// image order makes "strictly greater than the root" exactly the stated tie rule (value desc,
// image index asc).  NaN is the largest value, -0.0 == +0.0 (common.cuh: ordered_key).
//
// The image axis may be split across CTAs (grid.y) for load balance on small K; every
// (split, column) writes its k survivors to the workspace and topk_finish_kernel sorts the
// splits*k candidates of a column (warp-level bitonic sort on 64-bit (key, ~index) words) and
// emits indices (and values gathered from A, so they carry the input's exact bits).
#include "common.cuh"

namespace mcd {

constexpr int kScanUnroll = 8;

template <int COLS, int CAPTOT, int ROWS, int NSTAGE>
struct ScanCfg {
    static constexpr int kThreads = COLS + 32;
    static constexpr size_t kRingBytes = size_t(NSTAGE) * ROWS * COLS * 4;
    static constexpr size_t kHeapBytes = size_t(2) * CAPTOT * COLS * 4;
    static constexpr size_t kSmemBytes = kRingBytes + kHeapBytes + size_t(2) * NSTAGE * 8;
};

template <int COLS>
__device__ __forceinline__ void heap_replace_root(uint32_t *hi, uint32_t *lo, int k, int tid,
                                                  unsigned long long e) {
    int pos = 0;
    while (true) {
        int c = 2 * pos + 1;
        if (c >= k) break;
        unsigned long long cv = pack_key(hi[c * COLS + tid], lo[c * COLS + tid]);
        if (c + 1 < k) {
            unsigned long long cw = pack_key(hi[(c + 1) * COLS + tid], lo[(c + 1) * COLS + tid]);
            if (cw < cv) {
                cv = cw;
                c = c + 1;
            }
        }
        if (e <= cv) break;
        hi[pos * COLS + tid] = static_cast<uint32_t>(cv >> 32);
        lo[pos * COLS + tid] = static_cast<uint32_t>(cv);
        pos = c;
    }
    hi[pos * COLS + tid] = static_cast<uint32_t>(e >> 32);
    lo[pos * COLS + tid] = static_cast<uint32_t>(e);
}

// Fold every lane's pending list (slots k .. k+cnt-1) into its heap (slots 0 .. k-1).
template <int COLS>
__device__ __forceinline__ void fold_pending(uint32_t *hi, uint32_t *lo, int k, int tid, int &cnt, float &tau) {
    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    for (int j = 0; j < maxcnt; ++j) {
        if (j < cnt) {
            unsigned long long e = pack_key(hi[(k + j) * COLS + tid], lo[(k + j) * COLS + tid]);
            unsigned long long root = pack_key(hi[tid], lo[tid]);
            if (e > root) heap_replace_root<COLS>(hi, lo, k, tid, e);
        }
    }
    cnt = 0;
    tau = key_to_threshold(hi[tid]);
}

template <int COLS, int CAPTOT, int ROWS, int NSTAGE>
__global__ void __launch_bounds__(COLS + 32, 1)
topk_scan_kernel(const float *__restrict__ A, int64_t lda, int64_t N, int64_t K, int k, int64_t rows_per_split,
                 unsigned long long *__restrict__ cand, int bulk_ok) {
    using Cfg = ScanCfg<COLS, CAPTOT, ROWS, NSTAGE>;
    static_assert(ROWS % kScanUnroll == 0 && COLS % 32 == 0, "tile shape");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint32_t *hi = reinterpret_cast<uint32_t *>(smem_raw + Cfg::kRingBytes);
    uint32_t *lo = hi + CAPTOT * COLS;
    uint64_t *full = reinterpret_cast<uint64_t *>(lo + CAPTOT * COLS);
    uint64_t *empty = full + NSTAGE;

    const int tid = threadIdx.x;
    const int64_t c0 = int64_t(blockIdx.x) * COLS;
    const int ncols = static_cast<int>(min(int64_t(COLS), K - c0));
    const int split = blockIdx.y;
    const int64_t row0 = int64_t(split) * rows_per_split;
    const int64_t nrows = min(N, row0 + rows_per_split) - row0;
    const int ntiles = nrows > 0 ? static_cast<int>((nrows + ROWS - 1) / ROWS) : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], COLS / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= COLS) {
        // ---------------- producer warp: stream tiles of A into the ring ----------------
        const int lane = tid - COLS;
        const uint64_t policy = l2_policy_evict_first();
        for (int t = 0; t < ntiles; ++t) {
            const int stage = t % NSTAGE;
            const int use = t / NSTAGE;
            if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1);
            const int rows_here = static_cast<int>(min(int64_t(ROWS), nrows - int64_t(t) * ROWS));
            float *dst = ring + size_t(stage) * ROWS * COLS;
            const float *src = A + (row0 + int64_t(t) * ROWS) * lda + c0;
            if (bulk_ok) {
                if (lane == 0) mbar_arrive_expect_tx(&full[stage], uint32_t(rows_here) * uint32_t(ncols) * 4u);
                __syncwarp();
                for (int r = lane; r < rows_here; r += 32)
                    bulk_g2s(dst + r * COLS, src + int64_t(r) * lda, uint32_t(ncols) * 4u, &full[stage], policy);
            } else {
                // layout not 16-byte friendly: element copies by the producer warp
                for (int r = 0; r < rows_here; ++r)
                    for (int c = lane; c < ncols; c += 32) dst[r * COLS + c] = __ldg(src + int64_t(r) * lda + c);
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
            }
        }
        return;
    }

    // -------------------- consumers: one thread per neuron column --------------------
    const bool active = tid < ncols;
    for (int s = 0; s < k; ++s) {
        hi[s * COLS + tid] = 0u;
        lo[s * COLS + tid] = 0u;
    }
    const int pending_cap = CAPTOT - k;
    float tau = __uint_as_float(0x7FC00000u);   // NaN: "!(v <= tau)" admits everything until the heap is full
    int cnt = 0;

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t % NSTAGE;
        mbar_wait(&full[stage], (t / NSTAGE) & 1);
        const float *tile = ring + size_t(stage) * ROWS * COLS + tid;
        const int rows_here = static_cast<int>(min(int64_t(ROWS), nrows - int64_t(t) * ROWS));
        const uint32_t base_row = static_cast<uint32_t>(row0 + int64_t(t) * ROWS);
        for (int r0 = 0; r0 < rows_here; r0 += kScanUnroll) {
            if (__any_sync(0xffffffffu, cnt > pending_cap - kScanUnroll)) fold_pending<COLS>(hi, lo, k, tid, cnt, tau);
            float v[kScanUnroll];
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) v[u] = tile[(r0 + u) * COLS];
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) {
                if (!(v[u] <= tau) && active && (r0 + u < rows_here)) {
                    hi[(k + cnt) * COLS + tid] = ordered_key(v[u]);
                    lo[(k + cnt) * COLS + tid] = ~(base_row + uint32_t(r0 + u));
                    ++cnt;
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[stage]);
    }
    fold_pending<COLS>(hi, lo, k, tid, cnt, tau);

    if (active) {
        unsigned long long *dst = cand + (int64_t(split) * k) * K + c0 + tid;
        for (int s = 0; s < k; ++s) dst[int64_t(s) * K] = pack_key(hi[s * COLS + tid], lo[s * COLS + tid]);
    }
}

// One warp per column: sort the splits*k candidates (descending 64-bit words) and emit the top k.
constexpr int kFinishWarps = 4;

__global__ void __launch_bounds__(kFinishWarps * 32)
topk_finish_kernel(const unsigned long long *__restrict__ cand, int M, int Mpad, int k, int64_t K,
                   const float *__restrict__ A, int64_t lda, int64_t *__restrict__ idx64,
                   int32_t *__restrict__ idx32, float *__restrict__ vals) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col = int64_t(blockIdx.x) * kFinishWarps + warp;
    if (col >= K) return;
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(smem_raw) + size_t(warp) * Mpad;
    for (int i = lane; i < Mpad; i += 32) buf[i] = i < M ? cand[int64_t(i) * K + col] : 0ull;
    __syncwarp();
    for (int size = 2; size <= Mpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (Mpad >> 1); t += 32) {
                const int i = 2 * t - (t & (stride - 1));
                const int j = i + stride;
                const bool desc = (i & size) == 0;
                const unsigned long long a = buf[i], b = buf[j];
                if ((a < b) == desc) {
                    buf[i] = b;
                    buf[j] = a;
                }
            }
            __syncwarp();
        }
    }
    for (int r = lane; r < k; r += 32) {
        const uint32_t row = ~static_cast<uint32_t>(buf[r]);
        const int64_t o = int64_t(r) * K + col;
        if (idx64) idx64[o] = static_cast<int64_t>(row);
        if (idx32) idx32[o] = static_cast<int32_t>(row);
        if (vals) vals[o] = A[int64_t(row) * lda + col];
    }
}

// ---- host side ----------------------------------------------------------------------------------
struct TopkPlan {
    int variant;       // 0: k<=48, 1: k<=112, 2: k<=240, 3: k<=496
    int cols;
    int rows;          // tile rows
    int splits;
    int64_t rows_per_split;
    int mpad;
};

static bool make_plan(int64_t N, int64_t K, int64_t k, TopkPlan *p) {
    if (k <= 48) { p->variant = 0; p->cols = 128; }
    else if (k <= 112) { p->variant = 1; p->cols = 128; }
    else if (k <= 240) { p->variant = 2; p->cols = 64; }
    else if (k <= 496) { p->variant = 3; p->cols = 32; }
    else return false;
    p->rows = 32;
    const int64_t ncb = ceil_div<int64_t>(K, p->cols);
    int64_t splits = tunable(kTopkSplits);
    if (splits <= 0) {
        const int64_t target = 4 * int64_t(num_sms());
        splits = ceil_div<int64_t>(target, ncb);
    }
    const int64_t min_rows = k * 4 > 256 ? k * 4 : 256;
    int64_t max_splits = N / min_rows;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits * k > 4096) splits = 4096 / k;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int64_t rps = ceil_div<int64_t>(N, splits);
    rps = ceil_div<int64_t>(rps, p->rows) * p->rows;
    splits = ceil_div<int64_t>(N, rps);
    p->splits = static_cast<int>(splits);
    p->rows_per_split = rps;
    int mpad = 1;
    while (mpad < splits * k) mpad <<= 1;
    p->mpad = mpad;
    return true;
}

template <int COLS, int CAPTOT, int ROWS, int NSTAGE>
static int launch_scan(const float *A, int64_t lda, int64_t N, int64_t K, int k, const TopkPlan &p,
                       unsigned long long *cand, int bulk_ok, cudaStream_t st) {
    using Cfg = ScanCfg<COLS, CAPTOT, ROWS, NSTAGE>;
    auto kern = topk_scan_kernel<COLS, CAPTOT, ROWS, NSTAGE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::kSmemBytes)) != cudaSuccess)
        return MCD_ERR_CUDA;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(K, COLS)), static_cast<unsigned>(p.splits));
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(A, lda, N, K, k, p.rows_per_split, cand, bulk_ok);
    return check_launch();
}

}  // namespace mcd

extern "C" size_t mcd_topk_cols_workspace_bytes(int64_t N, int64_t K, int64_t k) {
    mcd::TopkPlan p;
    if (N < 1 || K < 1 || k < 1 || k > N || !mcd::make_plan(N, K, k, &p)) return 0;
    return size_t(p.splits) * size_t(k) * size_t(K) * sizeof(unsigned long long);
}

extern "C" int mcd_topk_cols_f32(const float *A, int64_t lda, int64_t N, int64_t K, int64_t k, int64_t *idx64_out,
                                 int32_t *idx32_out, float *vals_out, void *workspace, size_t workspace_bytes,
                                 mcd_stream_t stream) {
    using namespace mcd;
    if (!A || N < 1 || K < 1 || k < 1 || k > N || lda < K || N >= 0xFFFFFFFFll) return MCD_ERR_INVALID_ARGUMENT;
    TopkPlan p;
    if (!make_plan(N, K, k, &p)) return MCD_ERR_UNSUPPORTED;
    const size_t need = size_t(p.splits) * size_t(k) * size_t(K) * sizeof(unsigned long long);
    if (!workspace || workspace_bytes < need) return MCD_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto *cand = static_cast<unsigned long long *>(workspace);
    const int bulk_ok = (lda % 4 == 0) && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    int rc;
    switch (p.variant) {
        case 0: rc = launch_scan<128, 64, 32, 3>(A, lda, N, K, int(k), p, cand, bulk_ok, st); break;
        case 1: rc = launch_scan<128, 128, 32, 5>(A, lda, N, K, int(k), p, cand, bulk_ok, st); break;
        case 2: rc = launch_scan<64, 256, 32, 8>(A, lda, N, K, int(k), p, cand, bulk_ok, st); break;
        default: rc = launch_scan<32, 512, 32, 8>(A, lda, N, K, int(k), p, cand, bulk_ok, st); break;
    }
    if (rc != MCD_OK) return rc;
    const size_t fsmem = size_t(kFinishWarps) * p.mpad * sizeof(unsigned long long);
    if (cudaFuncSetAttribute(topk_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fsmem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    const unsigned fgrid = static_cast<unsigned>(ceil_div<int64_t>(K, kFinishWarps));
    topk_finish_kernel<<<fgrid, kFinishWarps * 32, fsmem, st>>>(cand, p.splits * int(k), p.mpad, int(k), K, A, lda,
                                                                idx64_out, idx32_out, vals_out);
    return check_launch();
}
