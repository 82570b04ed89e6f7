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
//   filter_scan_kernel   persistent CTA per SM: one producer lane streams [32 rows x 128 columns] boxes of A (16 KB,
//                        512-byte row pieces) into a shared-memory ring with TMA (cp.async.bulk.tensor.2d, mbarrier
//                        completion, L2 evict-first), 8 consumer warps take WHOLE tiles off the ring by ticket (a
//                        shared-memory counter), so no warp ever waits for another warp: the only synchronisation is
//                        the per-stage full / empty mbarrier pair.  A lane owns 4 adjacent columns of the tile
//                        (thresholds in registers), compares 4 rows x 4 columns per step and appends passing elements
//                        with predicated stores to its lane-private bag in shared memory (no atomics, no votes, no
//                        divergence).  A bag is emptied when it could overflow and whenever the warp moves to another
//                        column block: one global atomic per entry reserves the slot in the column's list, 4 in flight
//                        per lane.  Work items are (128-column block, row chunk) pairs handed out by a global counter,
//                        ~64 items per SM, so the tail of the launch is ~1.5 % and a launch over any column range
//                        (one pipeline chunk, one GPU's shard) fills the machine.
//   topk_select_kernel   warp per column: the k-th largest (key, ~row) word of the list by range-adaptive 32-bin
//                        histogram rounds (each round narrows the 64-bit key range 32-fold; ~2-3 rounds until <= 32
//                        candidates remain, then a 32-lane bitonic sort), the k words >= it are compacted and sorted
//                        with the register bitonic network, indices / values are emitted.
#pragma once

namespace mcd {

constexpr int kFCols = 128;                          // columns per block = TMA box width (512 bytes)
constexpr int kFRows = 32;                           // rows per tile
constexpr int kFConsumers = 8;                       // consumer warps per CTA (+ 1 producer warp)
constexpr int kFThreads = 32 * (kFConsumers + 1);
constexpr int kFBagCap = 32;                         // slots of a lane-private bag
constexpr int kFBagStep = 16;                        // a 4-row step appends at most 16 entries per lane
constexpr int kFMaxStages = 12;
constexpr uint32_t kFTileBytes = kFCols * kFRows * 4;
constexpr uint32_t kFBagSlotBytes = 32 * 8;          // bag[slot][lane] of uint2
constexpr int kFMaxLaunches = 64;                    // item counters per call (one per scan launch / pipeline chunk)

struct FilterMeta {
    int col0, row0, nvalid, pad;                     // col0 < 0: no more tiles
};

// behind the ring (nstage tiles) in dynamic shared memory
struct FilterTail {
    uint2 bag[kFConsumers][kFBagCap][32];
    FilterMeta meta[kFMaxStages];
    uint64_t full[kFMaxStages], empty[kFMaxStages];
    int next;                                        // ticket counter: the next tile to be taken by a consumer warp
};
__host__ __device__ inline size_t filter_smem_bytes(int nstage) { return size_t(nstage) * kFTileBytes + sizeof(FilterTail); }

__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();   // a lost arrival must not hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}

struct FilterArgs {
    int64_t N, K;
    int col_block0, n_col_blocks;        // this launch covers column blocks [col_block0, col_block0 + n_col_blocks)
    int chunk_tiles, n_chunks;           // a work item = chunk_tiles consecutive tiles of one column block
    int nstage, cap;
    const float *tau;                    // [K] start thresholds
    int *cnt;                            // [K] survivors per column (may exceed cap: the excess is dropped and the column flagged)
    unsigned long long *lists;           // [K][cap] survivor words (ordered key << 32 | ~row)
    int *item_ctr;                       // work-item counter of this launch (zero on entry)
};

__global__ void __launch_bounds__(kFThreads, 1)
filter_scan_kernel(const __grid_constant__ CUtensorMap tmap, const FilterArgs a) {
    extern __shared__ __align__(1024) unsigned char fsm[];
    FilterTail &t = *reinterpret_cast<FilterTail *>(fsm + size_t(a.nstage) * kFTileBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring_addr = smem_u32(fsm);
    const uint32_t full_addr = smem_u32(&t.full[0]), empty_addr = smem_u32(&t.empty[0]);
    const int nstage = a.nstage;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_init(&t.full[i], 1);
            mbar_init(&t.empty[i], 1);
        }
        t.next = 0;
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kFConsumers) {
        // ---- producer: one lane turns work items into TMA tile loads ------------------------------------------------
        if (lane != 0) return;
        const uint64_t policy = l2_policy_evict_first();
        const int n_items = a.n_col_blocks * a.n_chunks;
        const int64_t tiles_total = (a.N + kFRows - 1) / kFRows;
        int stage = 0, use = 0;
        auto next_stage = [&]() {
            if (use > 0) mbar_wait_bounded(empty_addr + stage * 8, (use - 1) & 1);
        };
        auto advance = [&]() {
            if (++stage == nstage) {
                stage = 0;
                ++use;
            }
        };
        for (;;) {
            const int item = atomicAdd(a.item_ctr, 1);
            if (item >= n_items) break;
            // column-block-major order: the blocks of a launch complete roughly in order
            const int b = item / a.n_chunks, r = item - b * a.n_chunks;
            const int col0 = (a.col_block0 + b) * kFCols;
            const int64_t tile0 = int64_t(r) * a.chunk_tiles;
            const int ntile = static_cast<int>(min(int64_t(a.chunk_tiles), tiles_total - tile0));
            for (int i = 0; i < ntile; ++i) {
                next_stage();
                const int64_t row0 = (tile0 + i) * kFRows;
                t.meta[stage] = FilterMeta{col0, static_cast<int>(row0), static_cast<int>(min(int64_t(kFRows), a.N - row0)), 0};
                mbar_arrive_expect_tx_addr(full_addr + stage * 8, kFTileBytes);
                tma_tile_g2s(ring_addr + stage * kFTileBytes, &tmap, col0, static_cast<int>(row0), full_addr + stage * 8, policy);
                advance();
            }
        }
        // one end marker per consumer warp
        for (int c = 0; c < kFConsumers; ++c) {
            next_stage();
            t.meta[stage] = FilterMeta{-1, 0, 0, 0};
            mbar_arrive_addr(full_addr + stage * 8);
            advance();
        }
        return;
    }

    // ---- consumers: whole tiles by ticket ----------------------------------------------------------------------------
    const uint32_t bag_base = smem_u32(&t.bag[warp][0][lane]);
    const uint32_t bag_limit = bag_base + uint32_t(kFBagCap - kFBagStep) * kFBagSlotBytes;   // beyond: a step may overflow
    uint32_t bp = bag_base;
    int cur_col0 = -1;
    float4 tau4 = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    const int cap = a.cap;

    // empty the lane-private bag into the survivor lists of the lane's 4 columns: one atomic per entry reserves the slot
    auto flush = [&]() {
        const int c = static_cast<int>((bp - bag_base) / kFBagSlotBytes);
        const int mx = __reduce_max_sync(0xffffffffu, c);
        for (int i = 0; i < mx; i += 4) {
            uint2 e[4];
            int pos[4], col[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < c) {
                    e[u] = t.bag[warp][i + u][lane];
                    col[u] = cur_col0 + lane * 4 + int(e[u].y & 3u);
                    pos[u] = atomicAdd(a.cnt + col[u], 1);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < c && pos[u] < cap)
                    a.lists[int64_t(col[u]) * cap + pos[u]] = pack_key(ordered_key(__uint_as_float(e[u].x)), ~(e[u].y >> 2));
        }
        bp = bag_base;
    };

    for (;;) {
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(&t.next, 1);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        const int stage = ticket % nstage;
        mbar_wait_bounded(full_addr + stage * 8, uint32_t(ticket / nstage) & 1u);
        const FilterMeta m = t.meta[stage];
        if (m.col0 < 0) break;
        if (m.col0 != cur_col0) {
            if (cur_col0 >= 0) flush();              // bag entries name their column relative to the block
            cur_col0 = m.col0;
            const int64_t c = int64_t(m.col0) + lane * 4;
            tau4.x = c + 0 < a.K ? a.tau[c + 0] : INFINITY;      // columns past K (zero-filled by TMA) never pass
            tau4.y = c + 1 < a.K ? a.tau[c + 1] : INFINITY;
            tau4.z = c + 2 < a.K ? a.tau[c + 2] : INFINITY;
            tau4.w = c + 3 < a.K ? a.tau[c + 3] : INFINITY;
        }
        const uint32_t tile = ring_addr + uint32_t(stage) * kFTileBytes + uint32_t(lane) * 16u;
        const uint32_t rowcode = uint32_t(m.row0) << 2;          // entry word 1: row << 2 | column within the lane
        const bool whole = m.nvalid == kFRows;
#pragma unroll
        for (int s = 0; s < kFRows / 4; ++s) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = lds_v4(tile + uint32_t(4 * s + i) * (kFCols * 4));
            if (s == kFRows / 4 - 1) {
                // the whole tile is in registers: hand the stage back before the last step's appends
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(empty_addr + stage * 8);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t rc = rowcode + (uint32_t(4 * s + i) << 2);
                const bool ok = whole || (4 * s + i) < m.nvalid;           // rows past N are zero-filled, not data
                if (ok && !(v[i].x <= tau4.x)) { sts_v2(bp, __float_as_uint(v[i].x), rc); bp += kFBagSlotBytes; }
                if (ok && !(v[i].y <= tau4.y)) { sts_v2(bp, __float_as_uint(v[i].y), rc | 1u); bp += kFBagSlotBytes; }
                if (ok && !(v[i].z <= tau4.z)) { sts_v2(bp, __float_as_uint(v[i].z), rc | 2u); bp += kFBagSlotBytes; }
                if (ok && !(v[i].w <= tau4.w)) { sts_v2(bp, __float_as_uint(v[i].w), rc | 3u); bp += kFBagSlotBytes; }
            }
            if (__any_sync(0xffffffffu, bp > bag_limit)) flush();
        }
    }
    if (cur_col0 >= 0) flush();
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

// bitonic sort, descending, of 32 * PER words held as word i = PER * lane + slot (strides below PER exchange inside a
// lane, the others with shfl.xor)
template <int PER>
__device__ __forceinline__ void bitonic_desc_regs(unsigned long long (&v)[PER], int lane) {
    constexpr int TOTAL = 32 * PER;
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
                        const unsigned long long x = v[e], y = v[e + stride < PER ? e + stride : e];
                        const unsigned long long hi = x > y ? x : y, lo = x > y ? y : x;
                        v[e] = desc ? hi : lo;
                        v[e + stride < PER ? e + stride : e] = desc ? lo : hi;
                    }
                }
            }
        }
    }
}

constexpr int kSelWarps = 8;

struct SelSmem {
    uint32_t hist[kSelWarps][32];
    unsigned long long small[kSelWarps][32];
    int count[kSelWarps];
};

// flags[col] = k when the column was resolved here, 0 when it has to be redone exactly (list short of k or overflowed)
template <int PER>
__global__ void __launch_bounds__(kSelWarps * 32)
topk_select_kernel(const unsigned long long *__restrict__ lists, const int *__restrict__ cnt, int cap, int k,
                   int64_t col_first, int64_t col_end, int64_t K, const float *__restrict__ A, int64_t lda,
                   int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ vals,
                   int *__restrict__ flags) {
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
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int i = lane; i < n; i += 32) {
        const unsigned long long e = ent[i];
        lo = e < lo ? e : lo;
        hi = e > hi ? e : hi;
    }
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
        for (int i = lane; i < n; i += 32) {
            const unsigned long long e = ent[i];
            if (e >= lo && e <= hi) atomicAdd(&sm.hist[warp][(e - lo) >> shift], 1u);
        }
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
    for (int i = lane; i < n; i += 32) {
        const unsigned long long e = ent[i];
        if (e >= lo && e <= hi) sm.small[warp][atomicAdd(&sm.count[warp], 1)] = e;
    }
    __syncwarp();
    unsigned long long x[1] = {lane < m ? sm.small[warp][lane] : 0ull};
    bitonic_desc_regs<1>(x, lane);
    const unsigned long long kth = __shfl_sync(0xffffffffu, x[0], need - 1);
    // the k words >= kth, sorted
    __syncwarp();
    if (lane == 0) sm.count[warp] = 0;
    for (int i = lane; i < 32 * PER; i += 32) sortbuf[warp][i] = 0ull;      // 0 sorts below every real word
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const unsigned long long e = ent[i];
        if (e >= kth) sortbuf[warp][atomicAdd(&sm.count[warp], 1)] = e;
    }
    __syncwarp();
    unsigned long long v[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) v[e] = sortbuf[warp][lane * PER + e];
    bitonic_desc_regs<PER>(v, lane);
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

}  // namespace mcd
