// K2 -- per-neuron top-k over the probe-image axis (replaces torch.topk(A, dim=0, k),
// reference concept_vit/similarity.py:55 / :82 / :107).
//
// A is [N images, K neurons] row-major, so one neuron's activations are strided by K floats and the
// only coalesced way through A is "a row segment per warp".  The unit of work is ONE WARP (a 32-thread
// CTA) that owns 32 adjacent neuron columns (128-byte row segments) and one slice of the image axis;
// 7 such warps are resident per SM and none of them ever synchronises with another:
//
//   feed    the warp streams [32 rows x 32 cols] tiles of A into its private shared-memory ring with
//           TMA tensor-tile loads (cp.async.bulk.tensor.2d, mbarrier completion, L2 evict-first: A is
//           read exactly once); lane 0 re-arms a stage as soon as its rows are in registers.
//   scan    a lane reads 4 adjacent columns of a row with one 16-byte LDS (quarter-warp q takes rows
//           q, q+4, ..., q+28 of the tile) and compares them with the 4 column thresholds it keeps in
//           registers (threshold = value of the column's current k-th best).  The few elements that
//           beat their threshold are appended -- predicated stores, no atomics, no divergence -- to
//           the pending list private to (quarter-warp, column).
//   fold    when a pending list could overflow in the next tile, lane c folds the lists of column c
//           into the column's kept set (an unsorted two-level min structure of (key, ~index) words:
//           groups in L2-resident global memory, group minima in shared memory, the group holding
//           the overall minimum mirrored in registers) and publishes the new threshold.  Only this
//           warp's stream pauses.
//
// A warp's kept sets only change between its scan steps and it scans rows in order, so when a step is
// scanned every kept entry has a smaller image index than every element of the step: "strictly greater
// than the k-th best" is then exactly the stated total order (value desc, image index asc).  The fold
// compares full 64-bit (key, ~index) words, so it is independent of the order of pending entries.
// NaN is the largest value, -0.0 == +0.0 (common.cuh).
//
// The image axis may be split across warps (grid.y) so that the warp count fills whole waves of the
// resident-warp slots; each (split, column) writes its survivors to the workspace and a finish kernel
// sorts the splits*k candidates of a column (bitonic network, in registers up to 256 candidates) and
// emits indices / values.
//
// Around the scan (mcd_topk_cols_f32 at the end of the file):
//   start threshold   sample_tilemax_kernel + sample_select_kernel: per column a value that, with
//                     overwhelming probability, has at least k column elements above it; the scan
//                     starts from it instead of -inf, column groups that come up short are rescanned
//                     exactly (per-column fill counts, only_flagged pass).
//   short columns     topk_small_kernel: for N <= 16384 rows of an L2-resident matrix an exact radix
//                     select (thread-block cluster per column group, DSMEM histogram reduction)
//                     replaces the scan, which would be all start-up there.
#include <cuda.h>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "topk_api.cuh"

namespace mcd {

constexpr int kUnitCols = 32;                        // columns per scan warp (= per CTA)
constexpr int kTileRows = 64;                        // rows per TMA tile
constexpr int kSampleRows = 32;                      // rows per tile of the sample pass (its own, smaller tiles)
constexpr int kScanThreads = 32;
constexpr int kQuads = 4;                            // quarter-warps: quarter q takes rows q, q+4, ..., q+28 of a tile
constexpr int kListCap = 12;                         // pending slots per (quarter, column) list
constexpr int kStepGroups = 8;                       // row groups (of 4 rows) scanned between two fold checks
// (measured at c4: short lists -- kListCap 4, a check every 2 row groups -- buy 14 resident warps per SM with the image
// axis split in two, and lose: 2.89 ms against 2.73 ms.  With nothing to insert 7 warps per SM already stream at
// 5.96 TB/s; what is left above that floor is insert work, and a split adds ~50 % of it.)
constexpr int kPendCap = kQuads * kListCap;          // pending slots per column
constexpr int kRowsPerQuad = kTileRows / kQuads;     // a tile adds at most this many entries to a list
constexpr int kMaxStages = 8;
constexpr int kMaxGroups = 8;                        // kept set: at most 8 groups (their minima live in registers)
enum FeedMode { kFeedElements = 0, kFeedTensorTile = 2 };

// one TMA instruction per [kTileRows x 32] tile; out-of-range rows / columns are zero-filled
__device__ __forceinline__ void tma_tile_g2s(uint32_t smem_dst, const CUtensorMap *tmap, int x, int y, uint32_t bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// Shared memory of a scan warp (arrays are [slot][32], so the warp's 32 columns hit distinct banks):
//   ring   nstage x [kTileRows x 32] fp32 tiles of A (only nstage of the kMaxStages are allocated)
//   pend   kPendCap pending entries {value bits, row} per column; the list of quarter-warp q starts at slot
//          q * kListCap
//   tau    current threshold per column;  pcnt  entries per (quarter, column) list, published before a fold
// Global memory (workspace, one region per warp; touched only by the folds, it stays in L2 because A is
// streamed with evict-first):
//   kept   per column the kept set as a [slot][32] array of (ordered key, ~row) u32 pairs: G = ceil(k/GROUP) groups
//          of GROUP entries, UNSORTED; slots >= k of the last group hold all-ones and are never the minimum
struct ScanSmem {
    float ring[kMaxStages][kTileRows][kUnitCols];
    uint2 pend[kPendCap][kUnitCols];
    float tau[kUnitCols];
    int pcnt[kQuads][kUnitCols];
    uint32_t gmh[kMaxGroups][kUnitCols], gml[kMaxGroups][kUnitCols];   // per group: its minimum entry (key, ~row)
    uint8_t gms[kMaxGroups][kUnitCols];                                 // ... and the slot holding it
    uint64_t full[kMaxStages];
};
__host__ __device__ inline size_t scan_smem_bytes(int nstage) {
    return sizeof(ScanSmem) - size_t(kMaxStages - nstage) * kTileRows * kUnitCols * 4;
}
__host__ __device__ inline int kept_group_size(int k) { return k <= 128 ? 16 : (k <= 256 ? 32 : 64); }
__host__ __device__ inline int kept_slots(int k) {
    const int g = kept_group_size(k);
    return (k + g - 1) / g * g;
}

__device__ __forceinline__ bool key_gt(uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl) {
    return ah > bh || (ah == bh && al > bl);
}

// The kept set of one column (owned by one lane) is a two-level min structure instead of a heap: G groups of
// GROUP unsorted (key, ~row) entries in global memory (L2-resident), per group its minimum entry in shared memory,
// and a REGISTER copy of the group that holds the overall minimum (the column's k-th best so far).  Replacing the
// minimum then needs no load at all: the entry is overwritten in the register copy (and stored to global, fire and
// forget), the group's new minimum comes from registers, the new overall minimum from the <= 8 group minima; only
// if it lies in another group are that group's GROUP entries loaded (independent 8-byte loads, one L2 round trip,
// first needed at the next replacement).
template <int GROUP>
struct Kept {
    uint2 *ent;                 // this lane's column: entry [slot] is ent[slot * 32]  (x = ordered key, y = ~row)
    uint32_t *gmh, *gml;        // shared memory: group g's minimum is gmh[g * 32], gml[g * 32], slot gms[g * 32]
    uint8_t *gms;
    uint32_t root_hi, root_lo;  // current overall minimum (the column's k-th best so far)
    int rg, rs;                 // its group and slot within the group
    float floor_tau;            // threshold while the set still has empty slots (NaN = admit everything)
    uint32_t ch[GROUP], cl[GROUP];   // register copy of group rg (statically indexed only)

    int cnt;                    // entries so far; the set is "full" (root valid, threshold live) once cnt == k

    // the caller has written all-ones into the padding slots (>= k) of the last group: never a minimum
    __device__ __forceinline__ void init(uint2 *ent_, ScanSmem &s, int lane) {
        ent = ent_;
        gmh = &s.gmh[0][lane];
        gml = &s.gml[0][lane];
        gms = &s.gms[0][lane];
        root_hi = root_lo = 0u;
        rg = rs = 0;
        cnt = 0;
        for (int g = 0; g < kMaxGroups; ++g) {
            gmh[g * kUnitCols] = gml[g * kUnitCols] = 0xFFFFFFFFu;
            gms[g * kUnitCols] = 0;
        }
#pragma unroll
        for (int i = 0; i < GROUP; ++i) ch[i] = cl[i] = 0xFFFFFFFFu;
    }

    // The first k entries of a column need no minimum search: store at the next free slot and keep the group's
    // minimum up to date (entries are only added, so it is a running minimum).  The k-th entry makes the set full:
    // the overall minimum is the smallest group minimum, and its group is loaded into the register copy.
    __device__ __forceinline__ void fill(uint32_t eh, uint32_t el, int k) {
        const int g = cnt / GROUP, sl = cnt % GROUP;
        ent[cnt * kUnitCols] = make_uint2(eh, el);
        if (key_gt(gmh[g * kUnitCols], gml[g * kUnitCols], eh, el)) {
            gmh[g * kUnitCols] = eh;
            gml[g * kUnitCols] = el;
            gms[g * kUnitCols] = static_cast<uint8_t>(sl);
        }
        if (++cnt == k) {
            uint32_t bh = 0xFFFFFFFFu, bl = 0xFFFFFFFFu;
            int bg = 0;
#pragma unroll
            for (int gg = kMaxGroups - 1; gg >= 0; --gg) {
                const uint32_t ah = gmh[gg * kUnitCols], al = gml[gg * kUnitCols];
                if (!key_gt(ah, al, bh, bl)) {
                    bh = ah;
                    bl = al;
                    bg = gg;
                }
            }
            root_hi = bh;
            root_lo = bl;
            rg = bg;
            rs = gms[bg * kUnitCols];
            const uint2 *ng = ent + bg * (GROUP * kUnitCols);
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
                const uint2 t = ng[i * kUnitCols];
                ch[i] = t.x;
                cl[i] = t.y;
            }
        }
    }

    // precondition: (eh, el) > root.  Overwrites the root entry with (eh, el) and re-establishes the minima.
    __device__ __forceinline__ void replace_min(uint32_t eh, uint32_t el) {
        uint2 *g = ent + rg * (GROUP * kUnitCols);
        g[rs * kUnitCols] = make_uint2(eh, el);
#pragma unroll
        for (int i = 0; i < GROUP; ++i) {
            ch[i] = (i == rs) ? eh : ch[i];
            cl[i] = (i == rs) ? el : cl[i];
        }
        // minimum (key, ~row) of the group (entries are distinct: the row index is part of them)
        uint32_t mh = ch[0], ml = cl[0];
        int ms = 0;
#pragma unroll
        for (int i = 1; i < GROUP; ++i) {
            const bool less = key_gt(mh, ml, ch[i], cl[i]);
            mh = less ? ch[i] : mh;
            ml = less ? cl[i] : ml;
            ms = less ? i : ms;
        }
        gmh[rg * kUnitCols] = mh;
        gml[rg * kUnitCols] = ml;
        gms[rg * kUnitCols] = static_cast<uint8_t>(ms);
        // overall minimum over the group minima (the changed group's values are taken from registers)
        uint32_t bh = 0xFFFFFFFFu, bl = 0xFFFFFFFFu;
        int bg = 0;
        uint32_t xh[kMaxGroups], xl[kMaxGroups];
#pragma unroll
        for (int gg = 0; gg < kMaxGroups; ++gg) {
            xh[gg] = gmh[gg * kUnitCols];
            xl[gg] = gml[gg * kUnitCols];
        }
#pragma unroll
        for (int gg = kMaxGroups - 1; gg >= 0; --gg) {
            const uint32_t ah = (gg == rg) ? mh : xh[gg], al = (gg == rg) ? ml : xl[gg];
            if (!key_gt(ah, al, bh, bl)) {
                bh = ah;
                bl = al;
                bg = gg;
            }
        }
        const int bs = (bg == rg) ? ms : gms[bg * kUnitCols];
        if (bg != rg) {
            const uint2 *ng = ent + bg * (GROUP * kUnitCols);
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
                const uint2 t = ng[i * kUnitCols];
                ch[i] = t.x;
                cl[i] = t.y;
            }
        }
        root_hi = bh;
        root_lo = bl;
        rg = bg;
        rs = bs;
    }
};

// whole warp: lane c folds the kQuads pending lists of column c into the column's kept set and publishes
// the new threshold
template <int GROUP>
__device__ __forceinline__ void fold_pending(ScanSmem &s, Kept<GROUP> &kept, int lane, int k) {
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
            const uint2 raw = s.pend[q * kListCap + e][lane];
            const uint32_t ch = ordered_key(__uint_as_float(raw.x)), cl = ~raw.y;
            if (kept.cnt < k) kept.fill(ch, cl, k);
            else if (key_gt(ch, cl, kept.root_hi, kept.root_lo)) kept.replace_min(ch, cl);
        }
    }
    s.tau[lane] = kept.cnt < k ? kept.floor_tau : key_to_threshold(kept.root_hi);
}

constexpr uint32_t kSlotBytes = kUnitCols * 8;       // one pending slot row

template <int GROUP>
__device__ __forceinline__ void publish_and_fold(ScanSmem &s, Kept<GROUP> &kept, int lane, int k, int q, int colq,
                                                 uint32_t list_first, uint32_t &p0, uint32_t &p1, uint32_t &p2,
                                                 uint32_t &p3, float4 &tau4) {
    s.pcnt[q][colq + 0] = static_cast<int>((p0 - list_first) / kSlotBytes);
    s.pcnt[q][colq + 1] = static_cast<int>((p1 - list_first) / kSlotBytes);
    s.pcnt[q][colq + 2] = static_cast<int>((p2 - list_first) / kSlotBytes);
    s.pcnt[q][colq + 3] = static_cast<int>((p3 - list_first) / kSlotBytes);
    __syncwarp();
    fold_pending<GROUP>(s, kept, lane, k);
    __syncwarp();
    tau4 = *reinterpret_cast<const float4 *>(&s.tau[colq]);
    p0 = list_first; p1 = list_first + 8; p2 = list_first + 16; p3 = list_first + 24;
}

// One warp per CTA: 32 adjacent neuron columns x one slice of the image axis.  The warp feeds itself: lane 0
// re-arms a ring stage with the next TMA tile as soon as the warp has the stage's rows in registers, so a warp
// that is busy folding only pauses its own stream; the other warps resident on the SM keep HBM busy.
template <int GROUP>
__global__ void __launch_bounds__(kScanThreads)
topk_scan_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ A, int64_t lda, int64_t N,
                 int64_t K, int k, int nstage, int64_t rows_per_split, unsigned long long *__restrict__ cand,
                 uint32_t *__restrict__ kept_ws, int feed, const float *__restrict__ tau0, int *__restrict__ flags,
                 int only_flagged, int group0) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the ring is declared with kMaxStages but only nstage stages are allocated: everything behind it moves up
    ScanSmem &s = *reinterpret_cast<ScanSmem *>(smem_raw - size_t(kMaxStages - nstage) * kTileRows * kUnitCols * 4);
    float(*ring)[kTileRows][kUnitCols] = reinterpret_cast<float(*)[kTileRows][kUnitCols]>(smem_raw);
    constexpr uint32_t kTileBytes = kTileRows * kUnitCols * 4;

    const int lane = threadIdx.x;
    const int64_t c0 = (int64_t(blockIdx.x) + group0) * kUnitCols;      // group0: first column group of a range launch
    const int ncols = static_cast<int>(min(int64_t(kUnitCols), K - c0));
    // second pass of the pre-threshold scheme: only column groups with a column that collected fewer than k elements
    // (over all splits) above its start threshold are redone, exactly, without one
    if (only_flagged && !__any_sync(0xffffffffu, lane < ncols && flags[c0 + lane] < k)) return;
    const int split = blockIdx.y;
    const int64_t row0 = int64_t(split) * rows_per_split;
    const int nrows = static_cast<int>(min(N, row0 + rows_per_split) - row0);
    const int ntiles = nrows > 0 ? (nrows + kTileRows - 1) / kTileRows : 0;
    const int G = (k + GROUP - 1) / GROUP;
    const uint64_t policy = l2_policy_evict_first();
    const uint32_t ring_addr = smem_u32(smem_raw);
    const uint32_t full_addr = smem_u32(&s.full[0]);
    const int tx = static_cast<int>(c0), ty0 = static_cast<int>(row0);

    if (lane == 0) {
        for (int i = 0; i < nstage; ++i) mbar_init(&s.full[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    // prologue: start the stream before touching anything else
    if (feed == kFeedTensorTile && lane == 0) {
        for (int t = 0; t < nstage && t < ntiles; ++t) {
            mbar_arrive_expect_tx(&s.full[t], kTileBytes);
            tma_tile_g2s(ring_addr + t * kTileBytes, &tmap, tx, ty0 + t * kTileRows, full_addr + t * 8, policy);
        }
    }
    // kept sets start empty; the NaN threshold admits everything until a column has seen k elements
    const int nslots = G * GROUP;
    uint2 *kept_ent = reinterpret_cast<uint2 *>(kept_ws) +
                      (size_t(blockIdx.y) * gridDim.x + blockIdx.x) * size_t(nslots) * kUnitCols + lane;
    for (int slot = k; slot < nslots; ++slot) kept_ent[slot * kUnitCols] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    Kept<GROUP> kept;
    kept.init(kept_ent, s, lane);
    // Start threshold: NaN admits everything; with a pre-threshold (a value known -- with overwhelming
    // probability -- to have at least k elements of the column above it) the insert-heavy start of the scan
    // disappears.  If fewer than k elements turn out to beat it, the group is flagged and redone without it.
    kept.floor_tau = (tau0 != nullptr && !only_flagged && lane < ncols) ? tau0[c0 + lane] : __uint_as_float(0x7FC00000u);
    s.tau[lane] = lane < ncols ? kept.floor_tau : INFINITY;                  // columns past K never pass
    __syncwarp();

    // A lane reads 4 adjacent columns of one row with one LDS.128; quarter-warp q takes rows q, q+4, ..., q+28
    // of a tile.  Elements that beat their column's threshold are appended, without atomics or divergent
    // branches, to the pending list PRIVATE to (quarter q, column): one predicated store + add each.
    const int q = lane >> 3;
    const int colq = (lane & 7) * 4;                 // first of this lane's 4 columns
    float4 tau4 = *reinterpret_cast<const float4 *>(&s.tau[colq]);
    const uint32_t list_first = smem_u32(&s.pend[q * kListCap][colq]);
    const uint32_t list_limit = list_first + (kListCap - kStepGroups) * kSlotBytes;   // beyond: a step may overflow
    uint32_t p0 = list_first, p1 = list_first + 8, p2 = list_first + 16, p3 = list_first + 24;
    const uint32_t lane_off = uint32_t(q * kUnitCols + colq) * 4u;   // this lane's first element inside a tile

    int stage = 0, use = 0;
    uint32_t row_tile = static_cast<uint32_t>(row0) + q;       // row index of this lane's first row in the next tile
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t, row_tile += kTileRows) {
        const uint32_t tile_addr = ring_addr + stage * kTileBytes;
        if (feed == kFeedTensorTile) {
            mbar_wait_addr(full_addr + stage * 8, use & 1);
        } else {
            // base pointer / row pitch not 16-byte aligned: the warp copies the tile itself
            const int rows_here = min(kTileRows, nrows - t * kTileRows);
            for (int r = 0; r < rows_here; ++r)
                ring[stage][r][lane] = lane < ncols ? __ldg(A + (row0 + int64_t(t) * kTileRows + r) * lda + c0 + lane) : 0.f;
            __syncwarp();
        }
        const bool full_tile = (t + 1) * kTileRows <= nrows;
        const int rows_left = nrows - t * kTileRows - q;           // > r  <=>  this lane's row r of the tile exists
        float4 v[kRowsPerQuad];                                    // this lane's 8 rows x 4 columns of the tile
#pragma unroll
        for (int i = 0; i < kRowsPerQuad; ++i) v[i] = lds_v4(tile_addr + lane_off + uint32_t(kQuads * i) * (kUnitCols * 4));
        __syncwarp();                                    // every lane has the tile's rows in registers
        if (feed == kFeedTensorTile && lane == 0 && t + nstage < ntiles) {
            fence_proxy_async();                         // generic-proxy reads before the async-proxy refill
            mbar_arrive_expect_tx_addr(full_addr + stage * 8, kTileBytes);
            tma_tile_g2s(tile_addr, &tmap, tx, ty0 + (t + nstage) * kTileRows, full_addr + stage * 8, policy);
        }
        if (++stage == nstage) {
            stage = 0;
            ++use;
        }
        // ~0.35 of the 128 elements of a row group pass on average: one vote per row group, and (warp-uniformly)
        // nothing else for the groups nobody appends from.  Rows past the end of the last, partial tile (stale ring
        // contents) are masked inside the branch only.
        // all votes first (independent chains: compare x4 -> vote), then the rarely taken append blocks
        uint32_t hit[kRowsPerQuad];
#pragma unroll
        for (int i = 0; i < kRowsPerQuad; ++i)
            hit[i] = __ballot_sync(0xffffffffu, !(v[i].x <= tau4.x) | !(v[i].y <= tau4.y) | !(v[i].z <= tau4.z) | !(v[i].w <= tau4.w));
#pragma unroll
        for (int i = 0; i < kRowsPerQuad; ++i) {
            if (hit[i] != 0u) {
                // (a fold earlier in this tile may have raised tau4 since the vote: the test below uses the new one)
                bool px = !(v[i].x <= tau4.x), py = !(v[i].y <= tau4.y), pz = !(v[i].z <= tau4.z), pw = !(v[i].w <= tau4.w);
                if (!full_tile) {
                    const bool valid = kQuads * i < rows_left;
                    px &= valid; py &= valid; pz &= valid; pw &= valid;
                }
                const uint32_t row = row_tile + uint32_t(kQuads * i);
                if (px) { sts_v2(p0, __float_as_uint(v[i].x), row); p0 += kSlotBytes; }
                if (py) { sts_v2(p1, __float_as_uint(v[i].y), row); p1 += kSlotBytes; }
                if (pz) { sts_v2(p2, __float_as_uint(v[i].z), row); p2 += kSlotBytes; }
                if (pw) { sts_v2(p3, __float_as_uint(v[i].w), row); p3 += kSlotBytes; }
            }
            if ((i + 1) % kStepGroups == 0) {
                // fold as soon as the next step could overflow a list
                const bool want = (p0 > list_limit) | (p1 > list_limit + 8) | (p2 > list_limit + 16) | (p3 > list_limit + 24);
                if (__any_sync(0xffffffffu, want))
                    publish_and_fold<GROUP>(s, kept, lane, k, q, colq, list_first, p0, p1, p2, p3, tau4);
            }
        }
    }
    publish_and_fold<GROUP>(s, kept, lane, k, q, colq, list_first, p0, p1, p2, p3, tau4);
    if (flags != nullptr && tau0 != nullptr && !only_flagged && lane < ncols) {
        // how many entries this (split, column) collected above the start threshold
        const int have = kept.cnt;
        atomicAdd(&flags[c0 + lane], have);
    }
    if (cand != nullptr && lane < ncols) {
        unsigned long long *dst = cand + (int64_t(split) * k) * K + c0 + lane;
        for (int i = 0; i < k; ++i) {
            const uint2 e = kept_ent[i * kUnitCols];
            dst[int64_t(i) * K] = i < kept.cnt ? pack_key(e.x, e.y) : 0ull;      // 0 sorts below every real entry
        }
    }
}

// ---- start threshold from a row sample ---------------------------------------------------------------------------
// sample_tilemax_kernel: per column the maximum (as an ordered key, so NaN counts as the largest value) of each
// sampled 32-row tile -- one tile out of every `stride` (32 at c4: 1/32 of A).  Pure streaming, 16-byte loads.
// sample_select_kernel : thread per column, the j-th largest of those tile maxima (small unsorted buffer in shared
// memory).  It is <= the j-th largest element of the sample, so the Poisson bound of make_plan() applies to it.
constexpr int kSampleCols = 128, kSampleThreads = 256, kSelectThreads = 64, kSelectMaxJ = 64;

__global__ void __launch_bounds__(kSampleThreads)
sample_tilemax_kernel(const float *__restrict__ A, int64_t lda, int64_t K, int64_t tile_row_stride, int vec_ok,
                      uint32_t *__restrict__ tilemax /*[ntiles][K]*/) {
    pdl_enter();
    __shared__ uint32_t red[kSampleThreads / 32][kSampleCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = int64_t(blockIdx.x) * kSampleCols + lane * 4;
    const int64_t r0 = int64_t(blockIdx.y) * tile_row_stride;
    // the maximum in float arithmetic (fmaxf drops NaN) and a NaN flag beside it: 3 instructions per element instead of the
    // 8 of an ordered-key maximum; the key is formed once per column at the end (NaN counts as the largest value)
    float fm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    bool nan_seen[4] = {false, false, false, false};
#pragma unroll
    for (int r = warp; r < kSampleRows; r += kSampleThreads / 32) {
        const float *src = A + (r0 + r) * lda + c;
        float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (vec_ok && c + 3 < K) {
            const float4 q = ldg_nc_v4(src);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (c + e < K) v[e] = __ldg(src + e);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            fm[e] = fmaxf(fm[e], v[e]);
            nan_seen[e] |= v[e] != v[e];
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) red[warp][lane * 4 + e] = nan_seen[e] ? 0xFFFFFFFFu : ordered_key(fm[e]);
    __syncthreads();
    if (threadIdx.x < kSampleCols) {
        uint32_t best = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kSampleThreads / 32; ++w) best = max(best, red[w][threadIdx.x]);
        const int64_t col = int64_t(blockIdx.x) * kSampleCols + threadIdx.x;
        if (col < K) tilemax[int64_t(blockIdx.y) * K + col] = best;
    }
}

__global__ void __launch_bounds__(kSelectThreads)
sample_select_kernel(const uint32_t *__restrict__ tilemax, int ntiles, int64_t K, int j, float *__restrict__ tau) {
    pdl_enter();
    __shared__ uint32_t top[kSelectMaxJ][kSelectThreads];      // the j largest keys seen so far, unsorted
    const int64_t col = int64_t(blockIdx.x) * kSelectThreads + threadIdx.x;
    if (col >= K) return;
    for (int i = 0; i < j; ++i) top[i][threadIdx.x] = 0u;
    uint32_t low = 0u;
    int low_at = 0;
    auto offer = [&](uint32_t v) {
        if (v > low) {
            top[low_at][threadIdx.x] = v;
            low = 0xFFFFFFFFu;
            for (int i = 0; i < j; ++i) {
                const uint32_t x = top[i][threadIdx.x];
                if (x < low) {
                    low = x;
                    low_at = i;
                }
            }
        }
    };
    // the loads of a batch are independent (the scan of the small buffer is not): 8 L2 round trips overlap
    constexpr int kBatch = 8;
    int t = 0;
    for (; t + kBatch <= ntiles; t += kBatch) {
        uint32_t v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) v[u] = tilemax[int64_t(t + u) * K + col];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) offer(v[u]);
    }
    for (; t < ntiles; ++t) offer(tilemax[int64_t(t) * K + col]);
    tau[col] = key_to_threshold(low);      // NaN (admit everything) if fewer than j tiles or the j-th largest is NaN
}

// The same for up to 128 sampled tiles (every shape the plans produce up to N = 131072 rows): a CTA takes 32 adjacent
// columns -- the tile maxima come in as coalesced 128-byte rows through shared memory -- and each warp then serves 4 of
// them: a lane holds 4 tile maxima of the column, and the j-th largest key is built bit by bit -- the largest v with
// #{keys >= v} >= j -- with one warp-wide integer reduction per bit (32 x ~8 instructions instead of a serial scan of a
// j-entry buffer per key).
constexpr int kSelectWarps = 8, kSelectWarpTiles = 128, kSelectCtaCols = 32;

__global__ void __launch_bounds__(kSelectWarps * 32)
sample_select_warp_kernel(const uint32_t *__restrict__ tilemax, int ntiles, int64_t K, int j, float *__restrict__ tau) {
    pdl_enter();
    __shared__ uint32_t keys[kSelectCtaCols][kSelectWarpTiles + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t col0 = int64_t(blockIdx.x) * kSelectCtaCols;
    {   // all 16 row loads of a thread in flight before the first store
        uint32_t in[kSelectWarpTiles / kSelectWarps];
#pragma unroll
        for (int i = 0; i < kSelectWarpTiles / kSelectWarps; ++i) {
            const int t = warp + i * kSelectWarps;
            in[i] = (t < ntiles && col0 + lane < K) ? __ldg(tilemax + int64_t(t) * K + col0 + lane) : 0u;
        }
#pragma unroll
        for (int i = 0; i < kSelectWarpTiles / kSelectWarps; ++i) keys[lane][warp + i * kSelectWarps] = in[i];
    }
    __syncthreads();
    // the warp's four columns (warp, warp + 8, ...) side by side: four independent compare / reduce chains
    constexpr int kPerWarp = kSelectCtaCols / kSelectWarps;
    uint32_t key[kPerWarp][4], v[kPerWarp];
#pragma unroll
    for (int u = 0; u < kPerWarp; ++u) {
        v[u] = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) key[u][e] = keys[warp + u * kSelectWarps][lane + 32 * e];
    }
    // (the 10 lowest key bits are left at zero: a threshold up to 2^-13 relative below the exact j-th largest tile maximum
    // admits ~0.1 % more survivors and saves a third of the steps)
#pragma unroll 2
    for (int bit = 31; bit >= 10; --bit) {
#pragma unroll
        for (int u = 0; u < kPerWarp; ++u) {
            const uint32_t cand = v[u] | (1u << bit);
            const int mine = int(key[u][0] >= cand) + int(key[u][1] >= cand) + int(key[u][2] >= cand) + int(key[u][3] >= cand);
            if (__reduce_add_sync(0xFFFFFFFFu, mine) >= j) v[u] = cand;
        }
    }
#pragma unroll
    for (int u = 0; u < kPerWarp; ++u) {
        const int64_t col = col0 + warp + u * kSelectWarps;
        if (lane == 0 && col < K) tau[col] = key_to_threshold(v[u]);      // v == 0: fewer than j tiles -> NaN (admit everything)
    }
}

static int launch_sample_select(const uint32_t *tilemax, int ntiles, int64_t K, int j, float *tau, cudaStream_t st) {
    if (ntiles <= kSelectWarpTiles)
        launch_pdl((sample_select_warp_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(K, kSelectCtaCols))), dim3(kSelectWarps * 32), 0, st, tilemax, ntiles, K, j, tau);
    else
        launch_pdl((sample_select_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(K, kSelectThreads))), dim3(kSelectThreads), 0, st, tilemax, ntiles, K, j, tau);
    return check_launch();
}

// One warp per column: sort the splits*k candidates (descending 64-bit words) and emit the top k.
constexpr int kFinishWarps = 4;

__global__ void __launch_bounds__(kFinishWarps * 32)
topk_finish_kernel(const unsigned long long *__restrict__ cand, int M, int Mpad, int k, int64_t K,
                   const float *__restrict__ A, int64_t lda, int64_t *__restrict__ idx64,
                   int32_t *__restrict__ idx32, float *__restrict__ vals, int64_t col_first, int64_t col_end,
                   const int *__restrict__ redo_flags) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col = col_first + int64_t(blockIdx.x) * (blockDim.x >> 5) + warp;     // 1 - 4 warps per CTA (shared memory)
    if (col >= col_end) return;
    if (redo_flags && redo_flags[col] >= k) return;      // resolved by the select kernel: only flagged columns are redone
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

// The common cases (splits * k <= 256, e.g. the default k = 100 of soft_wpmi with one or two splits): the same
// bitonic network with the words in registers, word i = PER * lane + slot (PER = 4 or 8).  Strides below PER exchange
// inside a lane, the others with shfl.xor; no shared memory, no loops left after unrolling -- about a third of the
// instructions of the generic kernel.  A CTA is 8 warps = 8 adjacent columns, so its index stores fill whole 32-byte
// sectors.
constexpr int kFinishRegWarps = 8;

template <int PER>
__global__ void __launch_bounds__(kFinishRegWarps * 32)
topk_finish_reg_kernel(const unsigned long long *__restrict__ cand, int M, int k, int64_t K,
                       const float *__restrict__ A, int64_t lda, int64_t *__restrict__ idx64,
                       int32_t *__restrict__ idx32, float *__restrict__ vals, int64_t col_first, int64_t col_end,
                       const int *__restrict__ redo_flags) {
    pdl_enter();
    constexpr int TOTAL = 32 * PER;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col = col_first + int64_t(blockIdx.x) * kFinishRegWarps + warp;
    if (col >= col_end) return;
    if (redo_flags && redo_flags[col] >= k) return;
    unsigned long long v[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int i = lane * PER + e;
        v[e] = i < M ? cand[int64_t(i) * K + col] : 0ull;
    }
#pragma unroll
    for (int size = 2; size <= TOTAL; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= PER) {
                // descending block <=> (i & size) == 0; the lower index of a pair keeps the maximum there
                const bool desc = size >= TOTAL || ((lane * PER) & size) == 0;
                const bool upper = (lane & (stride / PER)) != 0;
                const bool keep_max = desc != upper;
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[e], stride / PER);
                    const unsigned long long hi = v[e] > o ? v[e] : o, lo = v[e] > o ? o : v[e];
                    v[e] = keep_max ? hi : lo;
                }
            } else {
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    if ((e & stride) == 0) {
                        const bool desc = size >= TOTAL || ((lane * PER + e) & size) == 0;
                        const unsigned long long x = v[e], y = v[e + stride];
                        const unsigned long long hi = x > y ? x : y, lo = x > y ? y : x;
                        v[e] = desc ? hi : lo;
                        v[e + stride] = desc ? lo : hi;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int r = lane * PER + e;
        if (r < k) {
            const uint32_t row = ~static_cast<uint32_t>(v[e]);
            const int64_t o = int64_t(r) * K + col;
            if (idx64) idx64[o] = static_cast<int64_t>(row);
            if (idx32) idx32[o] = static_cast<int32_t>(row);
            if (vals) vals[o] = A[int64_t(row) * lda + col];
        }
    }
}

// ---- short columns: exact radix select -----------------------------------------------------------------------------
// With N up to a few thousand rows (the reference's real probe sets: 2 000 - 10 000 images, 24 - 768 neurons per layer)
// the streaming scan is all start-up: a warp spends its time filling and refilling kept sets, one L2-latency-bound
// replacement after the other (0.3 - 0.5 ms for 16 - 30 MB of input).  Such a matrix is L2-resident, so reading it a
// few times is cheap: a CTA of 32 warps takes 32 adjacent columns (lane = column, coalesced 128-byte row segments),
// the warps split the rows, and the k-th largest ordered key of every column is found by four 8-bit radix passes over
// per-column histograms in shared memory (lane-private banks: conflict-free shared atomics).  A fifth pass counts the
// elements equal to the k-th key per row slice -- the stated order takes the lowest image indices among equal values,
// so slice w may take min(its ties, what is still needed) -- and a sixth writes the k selected (key, ~row) words to
// the candidate array that the finish kernels sort.  Any base pointer / pitch (plain 4-byte loads).
constexpr int kSmallWarps = 32;
constexpr int kSmallMaxRows = 16384;
constexpr int64_t kSmallMaxBytes = int64_t(192) << 20;     // (mostly) L2-resident: the six passes re-read it

struct SmallSmem {
    uint32_t hist[256][kUnitCols];
    uint32_t ties[kSmallWarps][kUnitCols];      // ties in slice w, then (in place) how many of them slice w takes
    uint32_t prefix[kUnitCols], need[kUnitCols], outpos[kUnitCols];
    uint32_t cta_ties[kUnitCols], cta_quota[kUnitCols];
};

// A thread-block CLUSTER of R CTAs shares one group of 32 columns: CTA `rank` takes rows [rank*rpc, (rank+1)*rpc) with
// its 32 warps, so that a 768-neuron layer (24 column groups) still covers the machine (R = 4: 96 CTAs).  After each
// counting pass rank 0 adds the other CTAs' histograms to its own through distributed shared memory, walks the bins
// and the others fetch the new prefix; the tie quotas go rank by rank (lowest image indices first) and the output
// slots come from rank 0's counter (DSMEM atomics, k of them per column).
__global__ void __launch_bounds__(kSmallWarps * 32)
topk_small_kernel(const float *__restrict__ A, int64_t lda, int N, int64_t K, int k,
                  unsigned long long *__restrict__ cand) {
    pdl_enter();
    __shared__ SmallSmem s;
    const unsigned R = cluster_nctarank(), rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col = int64_t(blockIdx.x / R) * kUnitCols + lane;
    const bool active = col < K;
    const int rpc = (N + int(R) - 1) / int(R);                       // rows per CTA
    const int c0 = min(N, int(rank) * rpc), c1 = min(N, c0 + rpc);
    const int rps = (c1 - c0 + kSmallWarps - 1) / kSmallWarps;       // rows per warp
    const int r0 = min(c1, c0 + warp * rps), r1 = min(c1, r0 + rps);
    const float *src = A + (active ? col : 0);
    const uint32_t s_base = smem_u32(&s);
    // address of a field of rank q's SmallSmem
    auto remote = [&](const void *field, unsigned q) { return dsmem_addr(smem_u32(field), q); };
    (void)s_base;
    if (warp == 0) {
        s.prefix[lane] = 0u;
        s.need[lane] = static_cast<uint32_t>(k);
        s.outpos[lane] = 0u;
    }
    constexpr int kU = 16;      // independent loads in flight per lane
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256 * kUnitCols; i += kSmallWarps * 32) (&s.hist[0][0])[i] = 0u;
        __syncthreads();
        const uint32_t pre = s.prefix[lane];
        const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        if (active) {
            int r = r0;
            for (; r + kU <= r1; r += kU) {
                float v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) v[u] = __ldg(src + int64_t(r + u) * lda);
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t key = ordered_key(v[u]);
                    if ((key & himask) == pre) atomicAdd(&s.hist[(key >> shift) & 255u][lane], 1u);
                }
            }
            for (; r < r1; ++r) {
                const uint32_t key = ordered_key(__ldg(src + int64_t(r) * lda));
                if ((key & himask) == pre) atomicAdd(&s.hist[(key >> shift) & 255u][lane], 1u);
            }
        }
        cluster_sync_all();                     // every CTA's histogram is complete
        if (R > 1) {
            // CTA `rank` adds up its share of the bins over all CTAs and puts the totals into rank 0's histogram
            // (only the owner of a share reads or writes it, so rank 0's own counts are not raced on)
            const int per = 256 * kUnitCols / int(R);
            for (int i = int(rank) * per + threadIdx.x; i < int(rank + 1) * per; i += kSmallWarps * 32) {
                uint32_t *mine = &s.hist[0][0] + i;
                uint32_t part[8];
#pragma unroll
                for (unsigned q = 0; q < 8; ++q) part[q] = q < R ? ld_dsmem_u32(remote(mine, q)) : 0u;
                uint32_t sum = 0u;
#pragma unroll
                for (unsigned q = 0; q < 8; ++q) sum += part[q];
                st_dsmem_u32(remote(mine, 0), sum);
            }
            cluster_sync_all();                 // the totals are in rank 0
        }
        if (rank == 0) {
            if (warp == 0 && active) {
                // walk the bins from the largest digit down to the one that holds the `need`-th element (8 bins per
                // step so that the shared-memory loads overlap)
                uint32_t need = s.need[lane], acc = 0u;
                int b = 0;
                bool found = false;
                for (int hi8 = 255; hi8 >= 0 && !found; hi8 -= 8) {
                    uint32_t h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = s.hist[hi8 - j][lane];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (!found) {
                            if (acc + h[j] >= need) {
                                found = true;
                                b = hi8 - j;
                            } else {
                                acc += h[j];
                            }
                        }
                    }
                }
                s.prefix[lane] = pre | (uint32_t(b) << shift);
                s.need[lane] = need - acc;      // still to take among the elements that share the new prefix
            }
        }
        cluster_sync_all();                     // rank 0 has published prefix / need (and is done with the histograms)
        if (rank != 0 && warp == 0) {
            s.prefix[lane] = ld_dsmem_u32(remote(&s.prefix[lane], 0));
            s.need[lane] = ld_dsmem_u32(remote(&s.need[lane], 0));
        }
        __syncthreads();
    }
    const uint32_t T = s.prefix[lane];          // the column's k-th largest key
    // ties: how many elements equal to T lie in each row slice
    uint32_t t = 0u;
    if (active) {
        int r = r0;
        for (; r + kU <= r1; r += kU) {
            float v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) v[u] = __ldg(src + int64_t(r + u) * lda);
#pragma unroll
            for (int u = 0; u < kU; ++u) t += ordered_key(v[u]) == T;
        }
        for (; r < r1; ++r) t += ordered_key(__ldg(src + int64_t(r) * lda)) == T;
    }
    s.ties[warp][lane] = t;
    __syncthreads();
    if (warp == 0) {
        uint32_t tot = 0u;
        for (int w = 0; w < kSmallWarps; ++w) tot += s.ties[w][lane];
        s.cta_ties[lane] = tot;
    }
    cluster_sync_all();
    if (rank == 0 && warp == 0) {
        uint32_t left = s.need[lane];           // lowest image indices first: CTAs in rank order, then their slices
        for (unsigned q = 0; q < R; ++q) {
            const uint32_t have = q == 0 ? s.cta_ties[lane] : ld_dsmem_u32(remote(&s.cta_ties[lane], q));
            const uint32_t give = min(have, left);
            if (q == 0) s.cta_quota[lane] = give;
            else st_dsmem_u32(remote(&s.cta_quota[lane], q), give);
            left -= give;
        }
    }
    cluster_sync_all();
    if (warp == 0) {
        uint32_t left = s.cta_quota[lane];
        for (int w = 0; w < kSmallWarps; ++w) {
            const uint32_t q = min(s.ties[w][lane], left);
            s.ties[w][lane] = q;
            left -= q;
        }
    }
    __syncthreads();
    if (active) {
        uint32_t quota = s.ties[warp][lane];
        unsigned long long *dst = cand + col;
        const uint32_t counter = remote(&s.outpos[lane], 0);
        // in row order: the first `quota` ties of the slice are taken
        auto offer = [&](uint32_t key, int r) {
            bool take = key > T;
            if (key == T && quota > 0u) {
                take = true;
                --quota;
            }
            if (take) {
                const uint32_t pos = atom_add_dsmem_u32(counter, 1u);
                dst[int64_t(pos) * K] = pack_key(key, ~static_cast<uint32_t>(r));
            }
        };
        int r = r0;
        for (; r + kU <= r1; r += kU) {
            float v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) v[u] = __ldg(src + int64_t(r + u) * lda);
#pragma unroll
            for (int u = 0; u < kU; ++u) offer(ordered_key(v[u]), r + u);
        }
        for (; r < r1; ++r) offer(ordered_key(__ldg(src + int64_t(r) * lda)), r);
    }
    cluster_sync_all();                         // nobody leaves while its shared memory may still be addressed
}

}  // namespace mcd

#include "topk_filter.cuh"

namespace mcd {

// ---- host side ----------------------------------------------------------------------------------
constexpr size_t kSmemPerSM = 228 * 1024, kSmemCtaReserve = 1024;

constexpr int64_t kScanMaxK = 512;              // kept-set groups of the scan: 8 x 64 entries at most
constexpr int64_t kSelectMaxK = 16384;          // radix select + shared-memory sort of the k selected words (one warp: 128 KB)

static bool make_plan(int64_t N, int64_t K, int64_t k64, TopkPlan *p) {
    if (k64 < 1 || k64 > kScanMaxK) return false;
    const int k = static_cast<int>(k64);
    const int64_t sms = num_sms();
    // ring depth: enough scan warps per SM (the scan is latency-bound per warp) with a few tiles in flight each
    int nstage = static_cast<int>(tunable(kTopkStages));
    if (nstage < 2 || nstage > kMaxStages) nstage = 2;      // 2 x 8 KB tiles in flight per warp (measured: depth does not matter)
    p->nstage = nstage;
    p->smem = scan_smem_bytes(nstage);
    int occ = static_cast<int>(kSmemPerSM / (p->smem + kSmemCtaReserve));
    // registers: the kept set's root group and a 64-row tile live in registers (topk_scan_kernel<16|32|64>: 158 / 204 / 245)
    const int regs = k <= 128 ? 160 : (k <= 256 ? 208 : 248);
    if (occ > 65536 / (32 * regs)) occ = 65536 / (32 * regs);
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
    p->kept_bytes = size_t(ncb) * p->splits * size_t(kept_slots(k)) * kUnitCols * 8;   // (key, ~row) u32 pairs
    // Pre-threshold: the number of a column's top-k elements that fall into a 1/stride row sample is
    // ~ Poisson(k/stride), so the j = (lambda + 6 sqrt(lambda) + 4)-th largest sample value -- and a fortiori the j-th
    // largest of the sample's per-tile maxima, which is what is computed -- has >= k column elements above it
    // except with probability ~1e-8 (and a column that does come up short is redone exactly).
    p->pre_stride = 0;
    p->pre_k = 0;
    p->pre_rows = 0;
    p->pre_bytes = 0;
    if (tunable(kTopkPre) != 2 && N >= 8192 && k <= 256) {
        // sample one 32-row tile out of every `stride`; the estimate "j * stride elements above the threshold" needs
        // j well below the number of sampled tiles (correctness does not: the fill counts + exact redo guard it)
        static const int kStrides[] = {32, 16, 8, 4};
        for (int stride : kStrides) {
            const double lam = double(k) / stride;
            int pk = static_cast<int>(lam + 6.0 * sqrt(lam) + 4.0 + 0.999);
            if (pk < 8) pk = 8;
            const int64_t ntile = N / (int64_t(kSampleRows) * stride);
            if (pk <= kSelectMaxJ && 10 * ntile >= 11 * pk) {      // r = pk / ntile <= 0.91 (the list capacity carries m(r))
                p->pre_stride = stride;
                p->pre_k = pk;
                p->pre_rows = ntile * kSampleRows;
                p->pre_bytes = (size_t(K) * 4 + 255) / 256 * 256 + (size_t(K) * 4 + 255) / 256 * 256 +
                               size_t(ntile) * size_t(K) * 4;          // tau, fill counts, tile maxima
                break;
            }
        }
    }
    // Filter form (topk_filter.cuh) for long columns: survivor lists sized so that a column overflows with probability
    // <= 1e-10 (below); a column that does overflow, or comes up short of k, is redone exactly.
    p->filter = 0;
    p->f_cap = p->f_chunk_tiles = p->f_chunks = p->f_nstage = p->f_rows = 0;
    p->f_cnt_bytes = p->f_list_bytes = 0;
    // Taken whenever a row sample applies (N >= 8448 rows at k = 100): measured on B200 at K = 8192 .. 9216 against the
    // kept-set scan with the same start threshold: 0.25 / 0.67 ms at N = 10 000, 0.29 / 0.60 at 20 000, 0.33 / 0.74 at
    // 40 000; 0.10 / 0.81 ms at 20 000 x 512 (tools/time_k2_crossover.py).
    // The threshold is the j-th largest of `ntile` tile maxima: a fraction r = j / ntile of the tile maxima lies above it,
    // hence a fraction -ln(1 - r) / 32 of the ELEMENTS, i.e. stride * j * m(r) survivors per column with
    // m(r) = -ln(1 - r) / r (1.11 at c4, r = 18 / 97; 1.9 at N = 10 000, r = 60 / 78; -> 1 only for j << ntile).  The list
    // capacity below carries that factor.  (Without it the lists of short columns overflowed and the exact redo took
    // over -- which is what had made the filter form look slower than the kept-set scan below N = 40 000: 1.37 against
    // 0.62 ms at N = 20 000.  With it: 0.29 against 0.60 ms there, 0.25 against 0.67 ms at 10 000 x 9216.)
    const int64_t pre_ntile = p->pre_rows / kSampleRows;
    if (p->pre_stride > 0 && N < (int64_t(1) << 30) &&
        tunable(kTopkFilter) != 1 && tunable(kTopkVariant) == 0 &&
        tunable(kTopkSplits) <= 0) {
        // the number of column elements above the j-th largest of a 1/stride sample is ~ stride * Gamma(j): the list
        // holds stride * x elements with P(Gamma(j) > x) = P(Poisson(x) <= j - 1) <= 1e-10 (x = 59.25 for j = 18: 3.3 x the
        // mean; "mean + 7 sigma" of a normal fit would let one column in 200 000 overflow -- the tail is a gamma's)
        double x = double(p->pre_k);
        for (;; x += 0.25) {
            double term = exp(-x), cdf = 0.0;
            for (int i = 0; i < p->pre_k; ++i) {
                cdf += term;
                term *= x / double(i + 1);
            }
            if (cdf <= 1e-10 || x > 40.0 * p->pre_k) break;
        }
        double r = double(p->pre_k) / double(pre_ntile > p->pre_k ? pre_ntile : p->pre_k + 1);
        if (r > 0.92) r = 0.92;
        const double m_r = -log(1.0 - r) / r;            // survivors per column = stride * Gamma(j) * m(r)
        int cap = static_cast<int>(x * p->pre_stride * m_r) + 31;
        if (cap < 2 * k + 64) cap = 2 * k + 64;
        cap = cap / 32 * 32;
        const size_t list_bytes = size_t(K) * size_t(cap) * 8;
        if (list_bytes <= (size_t(16) << 30)) {
            p->filter = 1;
            p->f_cap = cap;
            int ns = static_cast<int>(tunable(kFilterStages));
            if (ns < 2 || ns > kFMaxStages) ns = 2;         // 18 KB per warp: 12 resident warps per SM, 96 KB in flight (measured best)
            p->f_nstage = ns;
            p->f_rows = tunable(kFilterOrder) == 16 ? 16 : kFRows;          // tunable "filter_order" = 16: 16-row tiles
            const int64_t tiles_total = N / p->f_rows, nblk = ceil_div<int64_t>(K, kFCols);
            int64_t ct = tunable(kFilterChunkTiles);
            if (ct <= 0) {
                // a work item = one CTA (scanner + drainer warp); ~80 items per resident CTA (11 per SM): short items keep
                // the last wave balanced (measured at c4: 31 tiles 2.19 ms, 62: 2.20, 85: 2.21, 124: 2.24 for the stage)
                ct = ceil_div<int64_t>(tiles_total * nblk, sms * 11 * 80);
                if (ct < 24) ct = 24;
            }
            if (ct > tiles_total) ct = tiles_total;
            if (ct < 1) ct = 1;
            p->f_chunk_tiles = static_cast<int>(ct);
            p->f_chunks = static_cast<int>(ceil_div<int64_t>(tiles_total, ct));
            p->f_cnt_bytes = ((size_t(K) + kFMaxLaunches) * 4 + 255) / 256 * 256;
            p->f_list_bytes = (list_bytes + 255) / 256 * 256;
        }
    }
    return true;
}

static size_t plan_pre_bytes_aligned(const TopkPlan &p) { return (p.pre_bytes + 255) / 256 * 256; }
static size_t plan_total_bytes(const TopkPlan &p) {
    return p.cand_bytes + p.kept_bytes + plan_pre_bytes_aligned(p) + p.f_cnt_bytes + p.f_list_bytes;
}
static bool takes_small_path(int64_t N, int64_t K) {
    return N <= kSmallMaxRows && (N <= 4096 || N * K * 4 <= kSmallMaxBytes) && tunable(kTopkSmall) != 1 &&
           tunable(kTopkSplits) <= 0 && tunable(kTopkVariant) != 1;
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

struct ScanArgs {
    const float *A;
    int64_t lda, N, K, rows_per_split;
    int k, feed;
    unsigned long long *cand;
    uint32_t *kept;
    const float *tau0;
    int *flags;
    int only_flagged;
    int group0;          // first 32-column group of a range launch (grid.x counts from it)
};

template <int GROUP>
static int launch_scan_t(dim3 grid, const TopkPlan &p, const CUtensorMap &map, const ScanArgs &a, cudaStream_t st) {
    auto kern = topk_scan_kernel<GROUP>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(p.smem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    launch_pdl((kern), dim3(grid), dim3(kScanThreads), p.smem, st, map, a.A, a.lda, a.N, a.K, a.k, p.nstage, a.rows_per_split, a.cand, a.kept,
                                             a.feed, a.tau0, a.flags, a.only_flagged, a.group0);
    return check_launch();
}

static int launch_scan(dim3 grid, const TopkPlan &p, const CUtensorMap &map, const ScanArgs &a, cudaStream_t st) {
    switch (kept_group_size(a.k)) {
        case 16: return launch_scan_t<16>(grid, p, map, a, st);
        case 32: return launch_scan_t<32>(grid, p, map, a, st);
        default: return launch_scan_t<64>(grid, p, map, a, st);
    }
}

static int launch_finish(const TopkPlan &p, const unsigned long long *cand, int k, int64_t K, const float *A, int64_t lda,
                         int64_t *idx64_out, int32_t *idx32_out, float *vals_out, cudaStream_t st, int64_t col_first = 0,
                         int64_t col_end = -1, const int *redo_flags = nullptr) {
    if (col_end < 0) col_end = K;
    const int64_t ncol = col_end - col_first;
    if (ncol <= 0) return MCD_OK;
    if (p.mpad <= 256 && tunable(kTopkVariant) != 2) {
        const unsigned fgrid = static_cast<unsigned>(ceil_div<int64_t>(ncol, kFinishRegWarps));
        if (p.mpad <= 128)
            launch_pdl((topk_finish_reg_kernel<4>), dim3(fgrid), dim3(kFinishRegWarps * 32), 0, st, cand, p.splits * k, k, K, A, lda, idx64_out, idx32_out,
                                                                              vals_out, col_first, col_end, redo_flags);
        else
            launch_pdl((topk_finish_reg_kernel<8>), dim3(fgrid), dim3(kFinishRegWarps * 32), 0, st, cand, p.splits * k, k, K, A, lda, idx64_out, idx32_out,
                                                                              vals_out, col_first, col_end, redo_flags);
        return check_launch();
    }
    // a warp sorts its column's mpad words in shared memory: as many warps per CTA as 160 KB hold (large k: one or two)
    int fw = static_cast<int>((size_t(160) << 10) / (size_t(p.mpad) * sizeof(unsigned long long)));
    if (fw > kFinishWarps) fw = kFinishWarps;
    if (fw < 1) return MCD_ERR_UNSUPPORTED;
    const size_t fsmem = size_t(fw) * p.mpad * sizeof(unsigned long long);
    if (cudaFuncSetAttribute(topk_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fsmem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    const unsigned fgrid = static_cast<unsigned>(ceil_div<int64_t>(ncol, fw));
    launch_pdl((topk_finish_kernel), dim3(fgrid), dim3(fw * 32), fsmem, st, cand, p.splits * k, p.mpad, k, K, A, lda, idx64_out, idx32_out,
                                                      vals_out, col_first, col_end, redo_flags);
    return check_launch();
}

// ---- filter form: prepare / begin / scan / finish (topk_api.cuh) ------------------------------------------------------
int topk_filter_prepare(const float *A, int64_t lda, int64_t N, int64_t K, int64_t k, void *workspace,
                        size_t workspace_bytes, TopkFilterCall *c) {
    if (!A || N < 1 || K < 1 || k < 1 || k > N || lda < K || N >= 0x7FFFFFFFll) return MCD_ERR_INVALID_ARGUMENT;
    TopkPlan p;
    if (!make_plan(N, K, k, &p)) return MCD_ERR_UNSUPPORTED;
    if (!p.filter || takes_small_path(N, K)) return MCD_ERR_UNSUPPORTED;
    const bool aligned = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    if (!aligned) return MCD_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < plan_total_bytes(p)) return MCD_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return MCD_ERR_INVALID_ARGUMENT;
    memset(&c->map_filter, 0, sizeof(CUtensorMap));
    memset(&c->map_scan, 0, sizeof(CUtensorMap));
    if (!make_tile_map(&c->map_filter, A, lda, N, K, kFCols, p.f_rows) || !make_tile_map(&c->map_scan, A, lda, N, K, kUnitCols, kTileRows))
        return MCD_ERR_UNSUPPORTED;
    c->plan = p;
    c->A = A;
    c->lda = lda;
    c->N = N;
    c->K = K;
    c->k = static_cast<int>(k);
    char *w = static_cast<char *>(workspace);
    c->cand = reinterpret_cast<unsigned long long *>(w);
    c->kept = reinterpret_cast<uint32_t *>(w + p.cand_bytes);
    char *pre = w + p.cand_bytes + p.kept_bytes;
    c->tau = reinterpret_cast<float *>(pre);
    c->flags = reinterpret_cast<int *>(pre + (size_t(K) * 4 + 255) / 256 * 256);
    c->tilemax = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(c->flags) + (size_t(K) * 4 + 255) / 256 * 256);
    c->cnt = reinterpret_cast<int *>(pre + plan_pre_bytes_aligned(p));
    c->lists = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(c->cnt) + p.f_cnt_bytes);
    return MCD_OK;
}

int topk_filter_begin(const TopkFilterCall &c, cudaStream_t st) {
    const TopkPlan &p = c.plan;
    if (cudaMemsetAsync(c.cnt, 0, (size_t(c.K) + kFMaxLaunches) * 4, st) != cudaSuccess) return MCD_ERR_CUDA;
    const int nsample = static_cast<int>(p.pre_rows / kSampleRows);
    dim3 sgrid(static_cast<unsigned>(ceil_div<int64_t>(c.K, kSampleCols)), static_cast<unsigned>(nsample));
    launch_pdl((sample_tilemax_kernel), dim3(sgrid), dim3(kSampleThreads), 0, st, c.A, c.lda, c.K, int64_t(kSampleRows) * p.pre_stride, 1, c.tilemax);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    return launch_sample_select(c.tilemax, nsample, c.K, p.pre_k, c.tau, st);
}

int topk_filter_scan(const TopkFilterCall &c, int64_t col0, int64_t col1, int launch_id, cudaStream_t st) {
    const TopkPlan &p = c.plan;
    if (col0 % 4 != 0 || col1 <= col0 || col1 > c.K || launch_id < 0 || launch_id >= kFMaxLaunches)
        return MCD_ERR_INVALID_ARGUMENT;
    FilterArgs a;
    a.N = c.N;
    a.K = c.K;
    a.col_begin = col0;
    a.col_end = col1;
    a.chunk_tiles = p.f_chunk_tiles;
    a.nstage = p.f_nstage;
    a.cap = p.f_cap;
    a.tau = c.tau;
    a.cnt = c.cnt;
    a.lists = c.lists;
    (void)launch_id;
    const int rows = p.f_rows;
    const size_t smem = filter_smem_bytes(p.f_nstage, rows);
    auto kern = rows == 16 ? filter_scan_kernel<16> : filter_scan_kernel<8>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) return MCD_ERR_CUDA;
    if (c.N >= rows) {
        // blockIdx.x = 128-column block (fastest: neighbours in a wave read neighbouring pieces of the same rows),
        // blockIdx.y = row chunk
        dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(col1 - col0, kFCols)), static_cast<unsigned>(p.f_chunks));
        launch_pdl((kern), dim3(grid), dim3(kFThreads), smem, st, c.map_filter, a);
    }
    int rc = check_launch();
    if (rc != MCD_OK || c.N % rows == 0) return rc;
    launch_pdl((filter_tail_rows_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(col1 - col0, 256))), dim3(256), 0, st, c.A, c.lda, c.N / rows * rows, c.N, col0, col1, c.tau, c.cnt, c.lists, p.f_cap);
    return check_launch();
}

int topk_filter_finish(const TopkFilterCall &c, int64_t col0, int64_t col1, int64_t *idx64, int32_t *idx32, float *vals,
                       cudaStream_t st) {
    const TopkPlan &p = c.plan;
    if (col0 % kUnitCols != 0 || col1 <= col0 || col1 > c.K) return MCD_ERR_INVALID_ARGUMENT;
    const unsigned sgrid = static_cast<unsigned>(ceil_div<int64_t>(col1 - col0, kSelWarps));
    int kpad = 32;
    while (kpad < c.k) kpad <<= 1;
#define MCD_SELECT(PER)                                                                                               \
    launch_pdl((topk_select_kernel<PER>), dim3(sgrid), dim3(kSelWarps * 32), 0, st, c.lists, c.cnt, p.f_cap, c.k, col0, col1, c.K, c.A, c.lda, \
                                                             idx64, idx32, vals, c.flags)
    switch (kpad) {
        case 32: MCD_SELECT(1); break;
        case 64: MCD_SELECT(2); break;
        case 128: MCD_SELECT(4); break;
        default: MCD_SELECT(8); break;          // make_plan: k <= 256 on this path
    }
#undef MCD_SELECT
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    // exact redo of the column groups with a flagged column (normally none: every CTA exits at once), then the
    // flagged columns' outputs from the redo's candidates
    const int64_t g0 = col0 / kUnitCols;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(col1 - col0, kUnitCols)), static_cast<unsigned>(p.splits));
    ScanArgs redo{c.A, c.lda, c.N, c.K, p.rows_per_split, c.k, kFeedTensorTile, c.cand, c.kept, nullptr, c.flags, 1,
                  static_cast<int>(g0)};
    rc = launch_scan(grid, p, c.map_scan, redo, st);
    if (rc != MCD_OK) return rc;
    return launch_finish(p, c.cand, c.k, c.K, c.A, c.lda, idx64, idx32, vals, st, col0, col1, c.flags);
}

}  // namespace mcd

extern "C" size_t mcd_topk_cols_workspace_bytes(int64_t N, int64_t K, int64_t k) {
    mcd::TopkPlan p;
    if (N < 1 || K < 1 || k < 1 || k > N) return 0;
    if (k > mcd::kScanMaxK) return k <= mcd::kSelectMaxK ? (size_t(k) * size_t(K) * 8 + 255) / 256 * 256 : 0;   // radix select: candidates only
    if (!mcd::make_plan(N, K, k, &p)) return 0;
    return plan_total_bytes(p);
}

extern "C" int mcd_topk_cols_f32(const float *A, int64_t lda, int64_t N, int64_t K, int64_t k, int64_t *idx64_out,
                                 int32_t *idx32_out, float *vals_out, void *workspace, size_t workspace_bytes,
                                 mcd_stream_t stream) {
    using namespace mcd;
    if (!A || N < 1 || K < 1 || k < 1 || k > N || lda < K || N >= 0x7FFFFFFFll) return MCD_ERR_INVALID_ARGUMENT;
    TopkPlan p;
    const bool large_k = k > kScanMaxK;         // beyond the kept sets of the scan (rank_reorder: k = 5 % of N): radix select
    if (large_k) {
        if (k > kSelectMaxK) return MCD_ERR_UNSUPPORTED;
        memset(&p, 0, sizeof(p));
        if (!workspace || workspace_bytes < size_t(k) * size_t(K) * 8) return MCD_ERR_WORKSPACE;
    } else {
        if (!make_plan(N, K, k, &p)) return MCD_ERR_UNSUPPORTED;
        if (!workspace || workspace_bytes < plan_total_bytes(p)) return MCD_ERR_WORKSPACE;
    }
    if (reinterpret_cast<uintptr_t>(workspace) % 8 != 0) return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto *cand = static_cast<unsigned long long *>(workspace);
    auto *kept = reinterpret_cast<uint32_t *>(static_cast<char *>(workspace) + p.cand_bytes);

    // short columns of an L2-resident matrix: exact radix select (tunable topk_small: 1 = never, else automatic)
    if (large_k || takes_small_path(N, K)) {
        // rows of a column group split over a cluster of R CTAs when the column groups alone would leave SMs idle
        const int64_t ncb_s = ceil_div<int64_t>(K, kUnitCols);
        int R = 1;
        while (R < 8 && ncb_s * R * 2 <= num_sms() && N / (R * 2) >= 512) R *= 2;
        if (tunable(kTopkSmall) >= 2 && tunable(kTopkSmall) <= 16) R = int(tunable(kTopkSmall)) / 2;   // test knob: 2,4,8,16 -> R = 1,2,4,8
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(static_cast<unsigned>(ncb_s * R));
        cfg.blockDim = dim3(kSmallWarps * 32);
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = static_cast<unsigned>(R);
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, topk_small_kernel, A, lda, int(N), K, int(k), cand) != cudaSuccess) {
            count_launch(1);
            return MCD_ERR_CUDA;
        }
        int rc0 = check_launch();
        if (rc0 != MCD_OK) return rc0;
        p.splits = 1;
        p.mpad = 1;
        while (p.mpad < int(k)) p.mpad <<= 1;
        return launch_finish(p, cand, int(k), K, A, lda, idx64_out, idx32_out, vals_out, st);
    }
    // feed: TMA tensor tiles need a 16-byte aligned base and row pitch; anything else takes element copies
    const bool aligned = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    int feed = aligned ? kFeedTensorTile : kFeedElements;
    if (tunable(kTopkVariant) == 1) feed = kFeedElements;        // test knob
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (feed == kFeedTensorTile && !make_tile_map(&map, A, lda, N, K, kUnitCols, kTileRows)) feed = kFeedElements;

    if (p.filter && feed == kFeedTensorTile && reinterpret_cast<uintptr_t>(workspace) % 256 == 0) {
        // long columns: sample -> filter scan -> select (+ exact redo of flagged groups)
        TopkFilterCall call;
        int frc = topk_filter_prepare(A, lda, N, K, k, workspace, workspace_bytes, &call);
        if (frc == MCD_OK) {
            frc = topk_filter_begin(call, st);
            if (frc != MCD_OK) return frc;
            frc = topk_filter_scan(call, 0, K, 0, st);
            if (frc != MCD_OK) return frc;
            return topk_filter_finish(call, 0, K, idx64_out, idx32_out, vals_out, st);
        }
        if (frc != MCD_ERR_UNSUPPORTED) return frc;
    }
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(K, kUnitCols)), static_cast<unsigned>(p.splits));
    ScanArgs main_args{A, lda, N, K, p.rows_per_split, int(k), feed, cand, kept, nullptr, nullptr, 0, 0};
    int rc;
    if (p.pre_stride > 0 && feed == kFeedTensorTile) {
        // pass 0: k'-th largest of a 1/32 row sample -> start threshold per column; pass 1: the real scan starting
        // from it; pass 2: exact redo (no start threshold) of the column groups that pass 1 flagged as short of k
        char *pre = static_cast<char *>(workspace) + p.cand_bytes + p.kept_bytes;
        float *tau = reinterpret_cast<float *>(pre);
        int *flags = reinterpret_cast<int *>(pre + (size_t(K) * 4 + 255) / 256 * 256);
        uint32_t *tilemax = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(flags) + (size_t(K) * 4 + 255) / 256 * 256);
        {
            // the sample is one 32-row tile out of every pre_stride tiles
            if (cudaMemsetAsync(flags, 0, size_t(K) * 4, st) != cudaSuccess) return MCD_ERR_CUDA;
            const int nsample = static_cast<int>(p.pre_rows / kSampleRows);
            dim3 sgrid(static_cast<unsigned>(ceil_div<int64_t>(K, kSampleCols)), static_cast<unsigned>(nsample));
            launch_pdl((sample_tilemax_kernel), dim3(sgrid), dim3(kSampleThreads), 0, st, A, lda, K, int64_t(kSampleRows) * p.pre_stride, 1, tilemax);
            rc = check_launch();
            if (rc != MCD_OK) return rc;
            rc = launch_sample_select(tilemax, nsample, K, p.pre_k, tau, st);
            if (rc != MCD_OK) return rc;
            main_args.tau0 = tau;
            main_args.flags = flags;
            rc = launch_scan(grid, p, map, main_args, st);
            if (rc != MCD_OK) return rc;
            main_args.only_flagged = 1;
        }
    }
    rc = launch_scan(grid, p, map, main_args, st);
    if (rc != MCD_OK) return rc;

    return launch_finish(p, cand, int(k), K, A, lda, idx64_out, idx32_out, vals_out, st);
}
