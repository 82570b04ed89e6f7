// K2, long columns -- the FILTER form of the per-neuron top-k (included by topk_cols.cu; replaces torch.topk(A, dim=0, k),
// reference concept_vit/similarity.py:55 / :82 / :107, for N >= 8192 probe images).
//
// The streaming pass over A does no selection at all.  Per column a start threshold is known (sample_tilemax_kernel +
// sample_select_kernel: with overwhelming probability at least k elements of the column exceed it, and about
// pre_k * pre_stride = 576 do at c4); the scan only FILTERS -- every element above its column's threshold is appended
// to that column's survivor list in global memory -- and a select kernel picks the exact top k of each list afterwards.
// A column whose list comes up short of k, or overflows its capacity, is flagged and redone exactly by
// topk_scan_kernel (only_flagged), so the result is exact whatever the data.
//
//   filter_scan_kernel   one WARP per CTA and per work item: 128 adjacent columns x a chunk of ~700 rows.  The warp feeds
//                        itself: [8 rows x 128 columns] tiles (4 KB, 512-byte row pieces) arrive in a private 4-stage
//                        shared-memory ring by TMA (cp.async.bulk.tensor.2d, mbarrier completion, L2 evict-first), and
//                        lane 0 re-arms a stage as soon as the warp holds its 8 rows in registers (8 x LDS.128 per
//                        lane) -- before any filtering.  8 such warps are resident per SM (128 KB in flight) and none
//                        ever waits for another, so a warp that is busy emptying its bag only pauses its own 1/8 of the
//                        SM's stream.  CTAs are dispatched in blockIdx order = adjacent column blocks of the same row
//                        chunk, so the warps that run at the same time read neighbouring 512-byte pieces of the same
//                        rows (DRAM pages are streamed contiguously).  A lane owns 4 adjacent columns (thresholds in
//                        registers).  Per row: four compares OR-ed into one predicate and a predicated append of the
//                        row's 4 values + the row index to the lane-private bag (no vote, no branch, no atomics).  The
//                        bag has two halves: emptying one only ISSUES four atomics (one per column, reserving the
//                        list slots) and switches to the other half; the entries are placed when that half is emptied
//                        in turn, long after the atomics have returned -- nobody waits for an L2 round trip.
//   topk_select_kernel   warp per column: the k-th largest (key, ~row) word of the list by range-adaptive 32-bin
//                        histogram rounds (each round narrows the 64-bit key range 32-fold; ~2-3 rounds until <= 32
//                        candidates remain, then a 32-lane bitonic sort), the k words >= it are compacted and sorted
//                        with the register bitonic network, indices / values are emitted.
#pragma once

namespace mcd {

constexpr int kFCols = 128;                          // columns per work item = TMA box width (a lane owns 4)
constexpr int kFRows = 8;                            // rows per tile (default; filter_scan_kernel<16> is the 16-row variant)
constexpr int kFThreads = 64;                        // two warps per CTA: the scanner and the drainer of its bags
constexpr int kFBagCap = 8;                          // entries of one half of a lane-private bag (an entry = 4 values of one row)
constexpr int kFBagStep = 4;                         // 4 rows append at most 4 entries per lane
constexpr int kFMaxStages = 8;
constexpr uint32_t kFRowBytes = kFCols * 4;
constexpr uint32_t kFBagValBytes = 32 * 16, kFBagRowBytes = 32 * 4;      // bag slot strides: float4 / row per lane
constexpr int kFMaxLaunches = 64;                    // (spare counters behind the survivor counts)

// behind the ring (nstage tiles) in dynamic shared memory
struct FilterTail {
    float4 bagv[2][kFBagCap][32];                    // two halves: the row's 4 values of a lane with a passing element
    uint32_t bagr[2][kFBagCap][32];                  // ... and the row
    uint64_t full[kFMaxStages];
    uint64_t ready[2], drained[2];                   // a half handed to the drainer warp / given back
    int pub_cnt[2][32];                              // entries per lane in a handed-over half
    int pub_last[2];                                 // 1: the item's last hand-over
};
__host__ __device__ inline size_t filter_smem_bytes(int nstage, int rows) { return size_t(nstage) * rows * kFRowBytes + sizeof(FilterTail); }

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// a lost arrival must not hang the GPU: after 2 s of waiting the kernel traps (the call then fails with a CUDA error)
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok, spins = 0;
    uint64_t t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
        if (!ok && (++spins & 255u) == 0u) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    } while (!ok);
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct FilterArgs {
    int64_t N, K;
    int64_t col_begin, col_end;          // this launch covers columns [col_begin, col_end): blockIdx.x counts 128-column blocks
    int chunk_tiles;                     // blockIdx.y counts row chunks of chunk_tiles whole 8-row tiles
    int nstage, cap;
    const float *tau;                    // [K] start thresholds
    int *cnt;                            // [K] survivors per column (may exceed cap: the excess is dropped and the column flagged)
    unsigned long long *lists;           // [K][cap] survivor words (ordered key << 32 | ~row)
};

template <int ROWS>
__global__ void __launch_bounds__(kFThreads)
filter_scan_kernel(const __grid_constant__ CUtensorMap tmap, const FilterArgs a) {
    pdl_enter();
    constexpr uint32_t kFTileBytes = ROWS * kFRowBytes;
    extern __shared__ __align__(1024) unsigned char fsm[];
    FilterTail &t = *reinterpret_cast<FilterTail *>(fsm + size_t(a.nstage) * kFTileBytes);
    const int lane = threadIdx.x & 31;
    const uint32_t ring_addr = smem_u32(fsm), full_addr = smem_u32(&t.full[0]);
    const int nstage = a.nstage;
    const int64_t c0 = a.col_begin + int64_t(blockIdx.x) * kFCols;
    const int64_t tiles_total = a.N / ROWS;          // whole tiles; filter_tail_rows_kernel takes the last N % ROWS rows
    const int64_t tile0 = int64_t(blockIdx.y) * a.chunk_tiles;
    const int ntile = static_cast<int>(min(int64_t(a.chunk_tiles), tiles_total - tile0));
    if (ntile <= 0) return;
    const uint64_t policy = l2_policy_evict_first();
    const int tx = static_cast<int>(c0), ty0 = static_cast<int>(tile0 * ROWS);

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) mbar_init(&t.full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t.ready[i], 1);
            mbar_init(&t.drained[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // prologue: start the stream before touching anything else
    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage && i < ntile; ++i) {
            mbar_arrive_expect_tx_addr(full_addr + i * 8, kFTileBytes);
            tma_tile_g2s(ring_addr + i * kFTileBytes, &tmap, tx, ty0 + i * ROWS, full_addr + i * 8, policy);
        }
    }
    // this lane's 4 columns and their thresholds (columns outside the launch's range never pass)
    const int64_t cur_col = c0 + lane * 4;
    float4 tau4;
    tau4.x = cur_col + 0 < a.col_end ? __ldg(a.tau + cur_col + 0) : INFINITY;
    tau4.y = cur_col + 1 < a.col_end ? __ldg(a.tau + cur_col + 1) : INFINITY;
    tau4.z = cur_col + 2 < a.col_end ? __ldg(a.tau + cur_col + 2) : INFINITY;
    tau4.w = cur_col + 3 < a.col_end ? __ldg(a.tau + cur_col + 3) : INFINITY;

    const int cap = a.cap;
    if (threadIdx.x >= 32) {
        // ---- drainer warp: empties the bag halves the scanner hands over (lane l takes lane l's entries): which of an
        // entry's four values passed is recomputed against tau, the lane's survivors are counted per column, ONE
        // atomicAdd per column with entries reserves the list slots, then the words go out.  Nothing here is on the
        // scanner's critical path: the stream never waits for an L2 round trip or for the scattered stores.
        unsigned long long *list = a.lists + cur_col * cap;
        int *cnt = a.cnt + cur_col;
#pragma unroll 1
        for (int s = 0;; ++s) {
            const int h = s & 1;
            mbar_wait_bounded(smem_u32(&t.ready[h]), uint32_t(s >> 1) & 1u);
            const int c = t.pub_cnt[h][lane];
            const int last = t.pub_last[h];
            const int mx = __reduce_max_sync(0xffffffffu, c);
            if (mx) {
                int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll 1
                for (int i = 0; i < mx; ++i) {
                    if (i < c) {
                        const float4 v = t.bagv[h][i][lane];
                        n0 += !(v.x <= tau4.x);
                        n1 += !(v.y <= tau4.y);
                        n2 += !(v.z <= tau4.z);
                        n3 += !(v.w <= tau4.w);
                    }
                }
                int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
                if (n0) o0 = atomicAdd(cnt + 0, n0);
                if (n1) o1 = atomicAdd(cnt + 1, n1);
                if (n2) o2 = atomicAdd(cnt + 2, n2);
                if (n3) o3 = atomicAdd(cnt + 3, n3);
                // One store instruction per entry slot for the whole warp (a second one only where an entry has two passing
                // values): the stream pays for every store INSTRUCTION the drainer issues, not for bytes or sectors
                // (profiles/r2_k2_history.md), so the four per-column stores of an entry are folded into one.
#pragma unroll 1
                for (int i = 0; i < mx; ++i) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    uint32_t nrow = 0u;
                    unsigned m = 0u;                 // which of the entry's four values passed
                    if (i < c) {
                        v = t.bagv[h][i][lane];
                        nrow = ~t.bagr[h][i][lane];
                        m = unsigned(!(v.x <= tau4.x)) | (unsigned(!(v.y <= tau4.y)) << 1) | (unsigned(!(v.z <= tau4.z)) << 2) |
                            (unsigned(!(v.w <= tau4.w)) << 3);
                    }
                    while (__any_sync(0xffffffffu, m != 0u)) {
                        if (m) {
                            const int b = __ffs(m) - 1;
                            m &= m - 1u;
                            const float x = b == 0 ? v.x : (b == 1 ? v.y : (b == 2 ? v.z : v.w));
                            const int pos = b == 0 ? o0 : (b == 1 ? o1 : (b == 2 ? o2 : o3));
                            o0 += b == 0;
                            o1 += b == 1;
                            o2 += b == 2;
                            o3 += b == 3;
                            if (pos < cap) list[int64_t(cap) * b + pos] = pack_key(ordered_key(x), nrow);
                        }
                    }
                }
            }
            __syncwarp();                            // every lane is done reading the half
            if (lane == 0) mbar_arrive(&t.drained[h]);
            if (last) break;
        }
        return;
    }

    // ---- scanner warp ----
    const uint32_t bagv0 = smem_u32(&t.bagv[0][0][lane]), bagr0 = smem_u32(&t.bagr[0][0][lane]);
    constexpr uint32_t kHalfV = kFBagCap * kFBagValBytes, kHalfR = kFBagCap * kFBagRowBytes;
    int half = 0, handed = 0;                        // handed = hand-overs so far; hand-over number s uses half s & 1
    uint32_t pv = bagv0, pr = bagr0;
    uint32_t half_limit = bagr0 + uint32_t(kFBagCap - kFBagStep) * kFBagRowBytes;

    // hand the current half to the drainer and continue in the other one (waiting, rarely, until it has been drained)
    auto hand_over = [&](int last) {
        t.pub_cnt[half][lane] = static_cast<int>((pr - (bagr0 + half * kHalfR)) / kFBagRowBytes);
        if (lane == 0) t.pub_last[half] = last;
        __syncwarp();                                // the bag entries and counts of all lanes, before the arrival
        if (lane == 0) mbar_arrive(&t.ready[half]);
        ++handed;
        half ^= 1;
        if (handed >= 2 && !last) mbar_wait_bounded(smem_u32(&t.drained[half]), uint32_t((handed - 2) >> 1) & 1u);
        pv = bagv0 + half * kHalfV;
        pr = bagr0 + half * kHalfR;
        half_limit = pr + uint32_t(kFBagCap - kFBagStep) * kFBagRowBytes;
    };
    // One row of the lane's slice: four compares OR-ed into one predicate, then the predicated append.  Written as one
    // asm block so that the predicate is consumed at once (left to itself the compiler hoists all 32 compares of a
    // tile in front of the appends and shuffles the predicates through a general register).
    auto append = [&](const float4 &v, uint32_t row0, int i) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
            "setp.gtu.f32 p, %2, %6;\n\t"
            "setp.gtu.or.f32 p, %3, %7, p;\n\t"
            "setp.gtu.or.f32 p, %4, %8, p;\n\t"
            "setp.gtu.or.f32 p, %5, %9, p;\n\t"
            "@p add.u32 r, %10, %11;\n\t"
            "@p st.shared.v4.f32 [%0], {%2,%3,%4,%5};\n\t"
            "@p st.shared.u32 [%1], r;\n\t"
            "@p add.u32 %0, %0, 512;\n\t"
            "@p add.u32 %1, %1, 128;\n\t}"
            : "+r"(pv), "+r"(pr)
            : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(tau4.x), "f"(tau4.y), "f"(tau4.z), "f"(tau4.w), "r"(row0), "r"(i)
            : "memory");
    };
    static_assert(kFBagValBytes == 512 && kFBagRowBytes == 128, "the append block hard-codes the bag strides");

    int stage = 0, use = 0;
    uint32_t row0 = static_cast<uint32_t>(ty0);
#pragma unroll 1
    for (int tl = 0; tl < ntile; ++tl, row0 += ROWS) {
        mbar_wait_bounded(full_addr + stage * 8, uint32_t(use) & 1u);
        const uint32_t tile = ring_addr + uint32_t(stage) * kFTileBytes + uint32_t(lane) * 16u;
        float4 v[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) v[i] = lds_v4(tile + uint32_t(i) * (kFCols * 4));
        __syncwarp();                                // every lane has the tile's rows in registers
        if (lane == 0 && tl + nstage < ntile) {
            fence_proxy_async();                     // generic-proxy reads before the async-proxy refill
            mbar_arrive_expect_tx_addr(full_addr + stage * 8, kFTileBytes);
            tma_tile_g2s(ring_addr + uint32_t(stage) * kFTileBytes, &tmap, tx, static_cast<int>(row0) + nstage * ROWS,
                         full_addr + stage * 8, policy);
        }
        if (++stage == nstage) {
            stage = 0;
            ++use;
        }
#pragma unroll 1
        for (int ph = 0; ph < ROWS / 4; ++ph) {
            // (a switch over static register indices: v[] must not be indexed dynamically)
            if (ph == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) append(v[i], row0, i);
            } else if (ph == 1) {
#pragma unroll
                for (int i = 4; i < 8; ++i) append(v[i], row0, i);
            } else if (ROWS > 8 && ph == 2) {
#pragma unroll
                for (int i = 8; i < 12; ++i) append(v[i < ROWS ? i : 0], row0, i);
            } else if (ROWS > 8) {
#pragma unroll
                for (int i = 12; i < 16; ++i) append(v[i < ROWS ? i : 0], row0, i);
            }
            // 4 rows add at most 4 entries per lane: switch halves when a lane has fewer than 4 free slots left
            if (__any_sync(0xffffffffu, pr > half_limit)) hand_over(0);
        }
    }
    hand_over(1);
}

// The last N % 8 rows of the matrix (the scan streams whole 8-row tiles only): thread per column.
__global__ void __launch_bounds__(256)
filter_tail_rows_kernel(const float *__restrict__ A, int64_t lda, int64_t row_begin, int64_t N, int64_t col_begin,
                        int64_t col_end, const float *__restrict__ tau, int *__restrict__ cnt,
                        unsigned long long *__restrict__ lists, int cap) {
    pdl_enter();
    const int64_t col = col_begin + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (col >= col_end) return;
    const float th = tau[col];
    for (int64_t r = row_begin; r < N; ++r) {
        const float v = A[r * lda + col];
        if (!(v <= th)) {
            const int pos = atomicAdd(cnt + col, 1);
            if (pos < cap) lists[col * cap + pos] = pack_key(ordered_key(v), ~static_cast<uint32_t>(r));
        }
    }
}

// ---- select ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long x = shfl_xor_u64(v, o);
        v = x < v ? x : v;
    }
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long x = shfl_xor_u64(v, o);
        v = x > v ? x : v;
    }
    return v;
}

constexpr int kSelWarps = 8;

struct SelSmem {
    uint32_t hist[kSelWarps][32];
    unsigned long long small[kSelWarps][32];
    int count[kSelWarps];
};

// flags[col] = k when the column was resolved here, 0 when it has to be redone exactly (list short of k or overflowed).
// The list (~576 words at c4, L2-resident: it was written moments ago) is read once per pass.  Keeping it in registers
// (96 registers, unrolled predicated passes) or in shared memory (30 KB per CTA) was measured and is slower: 0.21 ms and
// 0.19 ms against 0.13 ms at c4 -- the passes are instruction-bound, not load-bound.  Replacing the shared-memory atomics
// was measured too: ballot-prefix compaction instead of a counter 0.147 ms, a match.any leader per histogram bin 0.162 ms.
template <int PER>
__global__ void __launch_bounds__(kSelWarps * 32)
topk_select_kernel(const unsigned long long *__restrict__ lists, const int *__restrict__ cnt, int cap, int k,
                   int64_t col_first, int64_t col_end, int64_t K, const float *__restrict__ A, int64_t lda,
                   int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ vals,
                   int *__restrict__ flags) {
    pdl_enter();
    __shared__ SelSmem sm;
    __shared__ unsigned long long sortbuf[kSelWarps][32 * PER];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col = col_first + int64_t(blockIdx.x) * kSelWarps + warp;
    if (col >= col_end) return;
    const int n = cnt[col];
    if (n < k || n > cap) {
        if (lane == 0) flags[col] = 0;
        return;
    }
    if (lane == 0) flags[col] = k;
    const unsigned long long *ent = lists + col * cap;
    auto for_each = [&](auto &&fn) {
        for (int i = lane; i < n; i += 32) fn(ent[i]);
    };
    unsigned long long lo = ~0ull, hi = 0ull;
    for_each([&](unsigned long long w) {
        lo = w < lo ? w : lo;
        hi = w > hi ? w : hi;
    });
    lo = warp_min_u64(lo);
    hi = warp_max_u64(hi);
    // invariant: the need-th largest of the m words inside [lo, hi] is the k-th largest of the list
    int need = k, m = n;
    while (m > 32) {
        const unsigned long long range = hi - lo;             // > 0: m > 32 distinct words
        const int bits = 64 - __clzll(static_cast<long long>(range));
        const int shift = bits > 5 ? bits - 5 : 0;            // (range >> shift) < 32
        sm.hist[warp][lane] = 0u;
        __syncwarp();
        for_each([&](unsigned long long w) {
            if (w >= lo && w <= hi) atomicAdd(&sm.hist[warp][(w - lo) >> shift], 1u);
        });
        __syncwarp();
        const uint32_t h = sm.hist[warp][lane];
        uint32_t incl = h;                                    // words in bins >= lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += x;
        }
        const uint32_t above = incl - h;
        const unsigned hit = __ballot_sync(0xffffffffu, above < uint32_t(need) && uint32_t(need) <= incl);
        const int b = __ffs(hit) - 1;                         // exactly one bin holds the need-th largest
        need -= static_cast<int>(__shfl_sync(0xffffffffu, above, b));
        m = static_cast<int>(__shfl_sync(0xffffffffu, h, b));
        const unsigned long long nlo = lo + (static_cast<unsigned long long>(b) << shift);
        unsigned long long nhi = nlo + ((1ull << shift) - 1ull);
        if (nhi > hi || nhi < nlo) nhi = hi;
        lo = nlo;
        hi = nhi;
        __syncwarp();
    }
    // the <= 32 words left: sort them across the lanes and read off the need-th largest
    if (lane == 0) sm.count[warp] = 0;
    __syncwarp();
    for_each([&](unsigned long long w) {
        if (w >= lo && w <= hi) sm.small[warp][atomicAdd(&sm.count[warp], 1)] = w;
    });
    __syncwarp();
    unsigned long long x[1] = {lane < m ? sm.small[warp][lane] : 0ull};
    bitonic_desc_regs<1>(x, lane);
    const unsigned long long kth = __shfl_sync(0xffffffffu, x[0], need - 1);
    // the k words >= kth, sorted
    __syncwarp();
    if (lane == 0) sm.count[warp] = 0;
    for (int i = lane; i < 32 * PER; i += 32) sortbuf[warp][i] = 0ull;      // 0 sorts below every real word
    __syncwarp();
    for_each([&](unsigned long long w) {
        if (w >= kth) sortbuf[warp][atomicAdd(&sm.count[warp], 1)] = w;
    });
    __syncwarp();
    unsigned long long v[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) v[q] = sortbuf[warp][lane * PER + q];
    bitonic_desc_regs<PER>(v, lane);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int r = lane * PER + q;
        if (r < k) {
            const uint32_t row = ~static_cast<uint32_t>(v[q]);
            const int64_t o = int64_t(r) * K + col;
            if (idx64) idx64[o] = static_cast<int64_t>(row);
            if (idx32) idx32[o] = static_cast<int32_t>(row);
            if (vals) vals[o] = A[int64_t(row) * lda + col];
        }
    }
}

}  // namespace mcd
