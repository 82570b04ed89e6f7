// K2 -- per-neuron top-k over the probe-image axis (replaces torch.topk(A, dim=0, k),
// reference concept_vit/similarity.py:55 / :82 / :107).
//
// A is [N images, K neurons] row-major, so one neuron's activations are strided by K floats and the
// only coalesced way through A is "a row segment per warp".  The unit of work is ONE WARP (a 32-thread
// CTA) that owns 32 adjacent neuron columns (128-byte row segments) and one slice of the image axis;
// 6-8 such warps are resident per SM and none of them ever synchronises with another:
//
//   feed    the warp streams [32 rows x 32 cols] tiles of A into its private shared-memory ring with
//           TMA tensor-tile loads (cp.async.bulk.tensor.2d, mbarrier completion, L2 evict-first: A is
//           read exactly once); lane 0 re-arms a stage as soon as its rows are in registers.
//   scan    a lane reads 4 adjacent columns of a row with one 16-byte LDS (the 4 quarter-warps take 4
//           consecutive rows) and compares them with the 4 column thresholds it keeps in registers
//           (threshold = value of the column's current k-th best).  The few elements that beat their
//           threshold (~k ln(N/k) per column over the whole scan) are appended -- predicated stores,
//           no atomics, no divergence -- to the pending list private to (quarter-warp, column).
//   fold    when a pending list is nearly full, lane c folds the lists of column c into the column's
//           kept set (an unsorted two-level min structure of (key, ~index) words in L2-resident
//           global memory) and publishes the new threshold.  Only this warp's stream pauses.
//
// A warp's kept sets only change between its scan steps and it scans rows in order, so when a step is
// scanned every kept entry has a smaller image index than every element of the step: "strictly greater
// than the k-th best" is then exactly the stated total order (value desc, image index asc).  The fold
// compares full 64-bit (key, ~index) words, so it is independent of the order of pending entries.
// NaN is the largest value, -0.0 == +0.0 (common.cuh).
//
// The image axis may be split across warps (grid.y) so that the warp count fills whole waves of the
// resident-warp slots; each (split, column) writes its k survivors to the workspace and
// topk_finish_kernel sorts the splits*k candidates of a column (warp-level bitonic sort) and emits
// indices / values.
#include <cuda.h>
#include <cstring>

#include "common.cuh"

namespace mcd {

constexpr int kUnitCols = 32;                        // columns per scan warp (= per CTA)
constexpr int kTileRows = 32;                        // rows per TMA tile
constexpr int kSubRows = 16;                         // rows per scan step (4 quarter-warps x 4 rows)
constexpr int kScanThreads = 32;
constexpr int kQuads = 4;                            // quarter-warps: quarter q takes rows q, q+4, q+8, q+12 of a step
constexpr int kListCap = 8;                          // pending slots per (quarter, column) list
constexpr int kPendCap = kQuads * kListCap;          // pending slots per column
constexpr int kRowsPerQuad = kSubRows / kQuads;      // a step adds at most this many entries to a list
constexpr int kMaxStages = 8;
constexpr int kGroup = 16;                           // kept set: groups of kGroup entries + one minimum per group
enum FeedMode { kFeedElements = 0, kFeedTensorTile = 2 };

// one TMA instruction per [kTileRows x 32] tile; out-of-range rows / columns are zero-filled
__device__ __forceinline__ void tma_tile_g2s(void *smem_dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// Shared memory of a scan warp (arrays are [slot][32], so the warp's 32 columns hit distinct banks):
//   ring   nstage x [kTileRows x 32] fp32 tiles of A
//   pend   kPendCap pending entries (value bits, row) per column; the list of quarter-warp q starts at slot
//          q * kListCap
//   tau    current threshold per column;  pcnt  entries per (quarter, column) list, published before a fold
// Global memory (workspace, one region per warp, [slot][32] 64-bit words; touched only by the folds, it stays
// in L2 because A is streamed with evict-first):
//   kept   per column the kept set: G = ceil(k/kGroup) groups of kGroup (key, ~row) words, UNSORTED (slots
//          >= k of the last group hold the all-ones word and are never touched), then the G group minima
struct ScanSmem {
    float ring[kMaxStages][kTileRows][kUnitCols];
    unsigned long long pend[kPendCap][kUnitCols];
    float tau[kUnitCols];
    int pcnt[kQuads][kUnitCols];
    uint64_t full[kMaxStages];
};

__host__ __device__ inline int scan_groups(int k) { return (k + kGroup - 1) / kGroup; }
__host__ __device__ inline int kept_slots(int k) { return scan_groups(k) * (kGroup + 1); }
__host__ __device__ inline size_t scan_smem_bytes(int nstage) {
    return sizeof(ScanSmem) - size_t(kMaxStages - nstage) * kTileRows * kUnitCols * 4;
}

// The kept set of a column is a two-level min structure instead of a heap: replacing the current minimum
// (root, known to live in group rg) costs one pass over that group (find the slot holding root, take the
// group's new minimum) and one pass over the G group minima -- kGroup + G independent loads instead of a
// chain of log2(k) dependent ones.
__device__ __forceinline__ void kept_replace_min(unsigned long long *kept, int G, int col, unsigned long long e,
                                                 unsigned long long &root, int &rg) {
    unsigned long long *grp = kept + (rg * kGroup) * kUnitCols + col;
    unsigned long long x[kGroup];
#pragma unroll
    for (int i = 0; i < kGroup; ++i) x[i] = grp[i * kUnitCols];
    unsigned long long gmin = e;
    int hit = -1;
#pragma unroll
    for (int i = 0; i < kGroup; ++i) {
        const bool is_root = (hit < 0) && (x[i] == root);
        if (is_root) hit = i;
        else gmin = x[i] < gmin ? x[i] : gmin;
    }
    grp[hit * kUnitCols] = e;
    unsigned long long *gm = kept + (G * kGroup) * kUnitCols + col;
    gm[rg * kUnitCols] = gmin;
    unsigned long long best = gmin;
    int bg = rg;
    if (G <= kGroup) {
        // all group minima with independent loads (one L2 round trip), then the reduction
        unsigned long long m[kGroup];
#pragma unroll
        for (int g = 0; g < kGroup; ++g) m[g] = (g < G && g != rg) ? gm[g * kUnitCols] : ~0ull;
#pragma unroll
        for (int g = 0; g < kGroup; ++g)
            if (g != rg && (m[g] < best || (m[g] == best && g < bg))) {
                best = m[g];
                bg = g;
            }
    } else {
        for (int g = 0; g < G; ++g) {
            const unsigned long long m = (g == rg) ? gmin : gm[g * kUnitCols];
            if (m < best || (m == best && g < bg)) {
                best = m;
                bg = g;
            }
        }
    }
    root = best;
    rg = bg;
}

// whole warp: lane c folds the kQuads pending lists of column c into the column's kept set and publishes
// the new threshold
__device__ __forceinline__ void fold_pending(ScanSmem &s, unsigned long long *kept, int G, int lane,
                                             unsigned long long &root, int &rg) {
    int c[kQuads], total = 0;
#pragma unroll
    for (int q = 0; q < kQuads; ++q) {
        c[q] = s.pcnt[q][lane];
        total += c[q];
    }
    const int maxtotal = __reduce_max_sync(0xffffffffu, total);
    for (int j = 0; j < maxtotal; ++j) {
        if (j < total) {
            int q = 0, e = j;                        // j-th pending entry overall -> (list q, entry e)
#pragma unroll
            for (int qq = 0; qq < kQuads - 1; ++qq)
                if (q == qq && e >= c[qq]) {
                    e -= c[qq];
                    q = qq + 1;
                }
            const unsigned long long raw = s.pend[q * kListCap + e][lane];
            const float v = __uint_as_float(static_cast<uint32_t>(raw >> 32));
            const unsigned long long cand = pack_key(ordered_key(v), ~static_cast<uint32_t>(raw));
            if (cand > root) kept_replace_min(kept, G, lane, cand, root, rg);
        }
    }
    s.tau[lane] = key_to_threshold(static_cast<uint32_t>(root >> 32));
}

// One warp per CTA: 32 adjacent neuron columns x one slice of the image axis.  The warp feeds itself: lane 0
// re-arms a ring stage with the next TMA tile as soon as the warp has the stage's rows in registers, so a warp
// that is busy folding only pauses its own stream; the other warps resident on the SM keep HBM busy.
__global__ void __launch_bounds__(kScanThreads)
topk_scan_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ A, int64_t lda, int64_t N,
                 int64_t K, int k, int nstage, int64_t rows_per_split, unsigned long long *__restrict__ cand,
                 unsigned long long *__restrict__ kept_ws, int feed) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the ring is declared with kMaxStages but only nstage stages are allocated: everything behind it moves up
    ScanSmem &s = *reinterpret_cast<ScanSmem *>(smem_raw - size_t(kMaxStages - nstage) * kTileRows * kUnitCols * 4);
    float(*ring)[kTileRows][kUnitCols] = reinterpret_cast<float(*)[kTileRows][kUnitCols]>(smem_raw);

    const int lane = threadIdx.x;
    const int64_t c0 = int64_t(blockIdx.x) * kUnitCols;
    const int ncols = static_cast<int>(min(int64_t(kUnitCols), K - c0));
    const int split = blockIdx.y;
    const int64_t row0 = int64_t(split) * rows_per_split;
    const int nrows = static_cast<int>(min(N, row0 + rows_per_split) - row0);
    const int ntiles = nrows > 0 ? (nrows + kTileRows - 1) / kTileRows : 0;
    const int G = scan_groups(k);
    const uint64_t policy = l2_policy_evict_first();

    if (lane == 0) {
        for (int i = 0; i < nstage; ++i) mbar_init(&s.full[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    // prologue: start the stream before touching anything else
    if (feed == kFeedTensorTile && lane == 0) {
        for (int t = 0; t < nstage && t < ntiles; ++t) {
            mbar_arrive_expect_tx(&s.full[t], kTileRows * kUnitCols * 4u);
            tma_tile_g2s(&ring[t][0][0], &tmap, static_cast<int>(c0), static_cast<int>(row0 + int64_t(t) * kTileRows),
                         &s.full[t], policy);
        }
    }
    // kept sets start as k sentinels (word 0 sorts below every real entry); NaN threshold admits everything
    unsigned long long *kept = kept_ws + (size_t(blockIdx.y) * gridDim.x + blockIdx.x) * size_t(kept_slots(k)) * kUnitCols;
    for (int slot = 0; slot < kept_slots(k); ++slot)
        kept[slot * kUnitCols + lane] = (slot >= k && slot < G * kGroup) ? ~0ull : 0ull;   // padding: never the minimum
    s.tau[lane] = lane < ncols ? __uint_as_float(0x7FC00000u) : INFINITY;                 // columns past K never pass
    __syncwarp();

    // A lane reads 4 adjacent columns of one row with one LDS.128; quarter-warp q takes rows q, q+4, q+8, q+12
    // of a 16-row step.  Elements that beat their column's threshold are appended, without atomics or divergent
    // branches, to the pending list PRIVATE to (quarter q, column): one predicated store + add each.
    const int q = lane >> 3;
    const int colq = (lane & 7) * 4;                 // first of this lane's 4 columns
    float4 tau4 = *reinterpret_cast<const float4 *>(&s.tau[colq]);
    unsigned long long root = 0ull;                  // current minimum of the kept set of column `lane`, and its group
    int rg = 0;
    unsigned long long *const list_first = &s.pend[q * kListCap][colq];
    unsigned long long *const list_limit = list_first + (kListCap - kRowsPerQuad) * kUnitCols;   // beyond: a step may overflow
    unsigned long long *p0 = list_first, *p1 = list_first + 1, *p2 = list_first + 2, *p3 = list_first + 3;

    int stage = 0, use = 0;
    for (int t = 0; t < ntiles; ++t) {
        if (feed == kFeedTensorTile) {
            mbar_wait(&s.full[stage], use & 1);
        } else {
            // base pointer / row pitch not 16-byte aligned: the warp copies the tile itself
            const int rows_here = min(kTileRows, nrows - t * kTileRows);
            for (int r = 0; r < rows_here; ++r)
                ring[stage][r][lane] = lane < ncols ? __ldg(A + (row0 + int64_t(t) * kTileRows + r) * lda + c0 + lane) : 0.f;
            __syncwarp();
        }
        const int rows_tile = min(kTileRows, nrows - t * kTileRows);
        float4 v[kTileRows / kSubRows][kRowsPerQuad];
#pragma unroll
        for (int h = 0; h < kTileRows / kSubRows; ++h)
#pragma unroll
            for (int i = 0; i < kRowsPerQuad; ++i) {
                const int r = h * kSubRows + q + kQuads * i;
                v[h][i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                if (r < rows_tile) v[h][i] = *reinterpret_cast<const float4 *>(&ring[stage][r][colq]);
            }
        __syncwarp();                                    // every lane has its rows in registers
        if (feed == kFeedTensorTile && lane == 0 && t + nstage < ntiles) {
            fence_proxy_async();                         // generic-proxy reads before the async-proxy refill
            mbar_arrive_expect_tx(&s.full[stage], kTileRows * kUnitCols * 4u);
            tma_tile_g2s(&ring[stage][0][0], &tmap, static_cast<int>(c0),
                         static_cast<int>(row0 + int64_t(t + nstage) * kTileRows), &s.full[stage], policy);
        }
        if (++stage == nstage) {
            stage = 0;
            ++use;
        }
#pragma unroll
        for (int h = 0; h < kTileRows / kSubRows; ++h) {
            bool any = false;
#pragma unroll
            for (int i = 0; i < kRowsPerQuad; ++i)
                any |= !(v[h][i].x <= tau4.x) | !(v[h][i].y <= tau4.y) | !(v[h][i].z <= tau4.z) | !(v[h][i].w <= tau4.w);
            if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
            for (int i = 0; i < kRowsPerQuad; ++i) {
                const int r = h * kSubRows + q + kQuads * i;
                const bool valid = r < rows_tile;    // rows past the end read as -inf, which passes an unfilled (NaN) threshold
                const uint32_t row = static_cast<uint32_t>(row0) + uint32_t(t * kTileRows + r);
                if (valid && !(v[h][i].x <= tau4.x)) { *p0 = pack_key(__float_as_uint(v[h][i].x), row); p0 += kUnitCols; }
                if (valid && !(v[h][i].y <= tau4.y)) { *p1 = pack_key(__float_as_uint(v[h][i].y), row); p1 += kUnitCols; }
                if (valid && !(v[h][i].z <= tau4.z)) { *p2 = pack_key(__float_as_uint(v[h][i].z), row); p2 += kUnitCols; }
                if (valid && !(v[h][i].w <= tau4.w)) { *p3 = pack_key(__float_as_uint(v[h][i].w), row); p3 += kUnitCols; }
            }
            const bool want = (p0 > list_limit) | (p1 > list_limit + 1) | (p2 > list_limit + 2) | (p3 > list_limit + 3);
            if (__any_sync(0xffffffffu, want)) {
                s.pcnt[q][colq + 0] = static_cast<int>(p0 - list_first) / kUnitCols;
                s.pcnt[q][colq + 1] = static_cast<int>(p1 - list_first - 1) / kUnitCols;
                s.pcnt[q][colq + 2] = static_cast<int>(p2 - list_first - 2) / kUnitCols;
                s.pcnt[q][colq + 3] = static_cast<int>(p3 - list_first - 3) / kUnitCols;
                __syncwarp();
                fold_pending(s, kept, G, lane, root, rg);
                __syncwarp();
                tau4 = *reinterpret_cast<const float4 *>(&s.tau[colq]);
                p0 = list_first; p1 = list_first + 1; p2 = list_first + 2; p3 = list_first + 3;
            }
        }
    }
    s.pcnt[q][colq + 0] = static_cast<int>(p0 - list_first) / kUnitCols;
    s.pcnt[q][colq + 1] = static_cast<int>(p1 - list_first - 1) / kUnitCols;
    s.pcnt[q][colq + 2] = static_cast<int>(p2 - list_first - 2) / kUnitCols;
    s.pcnt[q][colq + 3] = static_cast<int>(p3 - list_first - 3) / kUnitCols;
    __syncwarp();
    fold_pending(s, kept, G, lane, root, rg);
    if (lane < ncols) {
        unsigned long long *dst = cand + (int64_t(split) * k) * K + c0 + lane;
        for (int i = 0; i < k; ++i) dst[int64_t(i) * K] = kept[i * kUnitCols + lane];
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
constexpr size_t kSmemPerSM = 228 * 1024, kSmemCtaReserve = 1024;

struct TopkPlan {
    int nstage, occ, splits, mpad;
    int64_t rows_per_split;
    size_t smem;
    size_t cand_bytes, kept_bytes;      // workspace: candidates [splits*k][K], then one kept region per scan warp
};

static bool make_plan(int64_t N, int64_t K, int64_t k64, TopkPlan *p) {
    if (k64 < 1 || k64 > 2048) return false;
    const int k = static_cast<int>(k64);
    const int64_t sms = num_sms();
    // ring depth: enough scan warps per SM (the scan is latency-bound per warp) with a few tiles in flight each
    int nstage = static_cast<int>(tunable(kTopkStages));
    if (nstage < 2 || nstage > kMaxStages) nstage = 4;
    p->nstage = nstage;
    p->smem = scan_smem_bytes(nstage);
    int occ = static_cast<int>(kSmemPerSM / (p->smem + kSmemCtaReserve));
    if (occ > 16) occ = 16;
    if (occ < 1) occ = 1;
    p->occ = occ;
    const int64_t min_rows = k * 4 > 256 ? k * 4 : 256;
    int64_t max_splits = N / min_rows;
    if (max_splits > 4096 / k) max_splits = 4096 / k;
    if (max_splits > 32) max_splits = 32;
    if (max_splits < 1) max_splits = 1;
    const int64_t ncb = ceil_div<int64_t>(K, kUnitCols);
    const int64_t slots = sms * occ;
    int64_t splits = tunable(kTopkSplits);
    if (splits <= 0) {
        // fewest splits of the image axis whose warp count fills whole waves of the resident-warp slots
        double best = -1e30;
        splits = 1;
        for (int64_t sp = 1; sp <= max_splits; ++sp) {
            const int64_t ctas = ncb * sp;
            const double eff = double(ctas) / double(ceil_div<int64_t>(ctas, slots) * slots);
            const double score = eff - 0.02 * double(sp - 1);   // every split re-fills its kept sets and adds merge work
            if (score > best) {
                best = score;
                splits = sp;
            }
        }
    }
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = ceil_div<int64_t>(N, splits);
    rps = ceil_div<int64_t>(rps, kTileRows) * kTileRows;
    p->splits = static_cast<int>(ceil_div<int64_t>(N, rps));
    p->rows_per_split = rps;
    int mpad = 1;
    while (mpad < p->splits * k) mpad <<= 1;
    p->mpad = mpad;
    p->cand_bytes = (size_t(p->splits) * size_t(k) * size_t(K) * 8 + 255) / 256 * 256;
    p->kept_bytes = size_t(ncb) * p->splits * size_t(kept_slots(k)) * kUnitCols * 8;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2-D tensor map over A [N rows, K cols] with a [rows x cols] box (no swizzle: lanes read 4 adjacent columns)
static bool make_tile_map(CUtensorMap *map, const float *A, int64_t lda, int64_t N, int64_t K, int cols, int rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(N)};
    cuuint64_t strides[1] = {cuuint64_t(lda) * sizeof(float)};
    cuuint32_t box[2] = {cuuint32_t(cols), cuuint32_t(rows)};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(A), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mcd

extern "C" size_t mcd_topk_cols_workspace_bytes(int64_t N, int64_t K, int64_t k) {
    mcd::TopkPlan p;
    if (N < 1 || K < 1 || k < 1 || k > N || !mcd::make_plan(N, K, k, &p)) return 0;
    return p.cand_bytes + p.kept_bytes;
}

extern "C" int mcd_topk_cols_f32(const float *A, int64_t lda, int64_t N, int64_t K, int64_t k, int64_t *idx64_out,
                                 int32_t *idx32_out, float *vals_out, void *workspace, size_t workspace_bytes,
                                 mcd_stream_t stream) {
    using namespace mcd;
    if (!A || N < 1 || K < 1 || k < 1 || k > N || lda < K || N >= 0x7FFFFFFFll) return MCD_ERR_INVALID_ARGUMENT;
    TopkPlan p;
    if (!make_plan(N, K, k, &p)) return MCD_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < p.cand_bytes + p.kept_bytes) return MCD_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 8 != 0) return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto *cand = static_cast<unsigned long long *>(workspace);
    auto *kept = reinterpret_cast<unsigned long long *>(static_cast<char *>(workspace) + p.cand_bytes);

    // feed: TMA tensor tiles need a 16-byte aligned base and row pitch; anything else takes element copies
    const bool aligned = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    int feed = aligned ? kFeedTensorTile : kFeedElements;
    if (tunable(kTopkVariant) == 1) feed = kFeedElements;        // test knob
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (feed == kFeedTensorTile && !make_tile_map(&map, A, lda, N, K, kUnitCols, kTileRows)) feed = kFeedElements;

    if (cudaFuncSetAttribute(topk_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(p.smem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(K, kUnitCols)), static_cast<unsigned>(p.splits));
    topk_scan_kernel<<<grid, kScanThreads, p.smem, st>>>(map, A, lda, N, K, int(k), p.nstage, p.rows_per_split, cand,
                                                         kept, feed);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;

    const size_t fsmem = size_t(kFinishWarps) * p.mpad * sizeof(unsigned long long);
    if (cudaFuncSetAttribute(topk_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fsmem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    const unsigned fgrid = static_cast<unsigned>(ceil_div<int64_t>(K, kFinishWarps));
    topk_finish_kernel<<<fgrid, kFinishWarps * 32, fsmem, st>>>(cand, p.splits * int(k), p.mpad, int(k), K, A, lda,
                                                                idx64_out, idx32_out, vals_out);
    return check_launch();
}
