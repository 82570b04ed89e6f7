// rank_reorder scoring (reference concept_vit/similarity.py:99-132), scope row f4.
//
// Per neuron j, over its top_n = int(0.05 N) probe images (K2 supplies indices and values):
//   x[r,c]    = P[idx[r,j], c]                       raw cosines of the selected images
//   rank[r,c] = position of x[r,c] among x[:,c] (0 = smallest; ties by r, i.e. a stable argsort)
//   err[c]    = mean_r |t[r] - asc[rank[r,c]]|^p / baseline_j,   t = the neuron's top_n activations (descending),
//               asc = t reversed
//   out[j,c]  = -err[c] / (mean_r x[r,c])^scale_p    (NaN when the mean cosine is negative, as in the reference)
// baseline_j = mean over 5 random permutations of |asc[r] - asc[perm[r]]|^p; the reference draws them with
// torch.randperm from the global CPU generator, so the host replays that stream and passes the permutations in.
//
#include "common.cuh"

namespace mcd {


__device__ __forceinline__ float abs_pow(float d, float p) {
    const float a = fabsf(d);
    if (p == 3.f) return a * a * a;
    if (p == 2.f) return a * a;
    if (p == 1.f) return a;
    return powf(a, p);
}

// ---- sort-based ranks (O(n log^2 n) per (neuron, concept) column instead of O(n^2) counting) -----------------------------
// Sorting the words (ordered cosine << 32 | row r) in DESCENDING order puts the element of ascending rank n-1-q at
// position q (ties: the larger r first, i.e. the lower r gets the lower ascending rank -- the stated stable order).
// With asc[i] = t[n-1-i] the reference's asc[rank[r]] is then t[q], so
//     err[c] = mean_q | t[row(q)] - t[q] |^p
// needs no scatter: walk the sorted positions.
//
// rank_sorted_warp_kernel<PER>: n <= 32 * PER (PER = 8: n <= 256, PER = 16: n <= 512).  CTA = 8 warps = 8 adjacent
// concepts of one neuron (the 8 gathers of a probe image share one 32-byte sector); a warp sorts its column in registers.
// The sorted words are 32 bits wide: the top 24 (23) bits of the ordered cosine and the row r in the low 8 (9) bits -- half
// the shuffles and a third of the compare instructions of a 64-bit sort.  Cosines that agree in those top bits (about one
// pair per column for real data) end up adjacent but possibly in the wrong order: such elements are detected after the
// sort (equal top bits with a neighbour) and their exact position in the full (ordered cosine, r) order is counted
// directly from the column kept in shared memory.  A constant column degenerates to counting, n^2 / 32 steps.
template <int PER>
__global__ void __launch_bounds__(256)
rank_sorted_warp_kernel(const float *__restrict__ P, int64_t ldp, int C, const int32_t *__restrict__ idx,
                        const float *__restrict__ vals, int64_t K, int n, float *__restrict__ den_out, float p,
                        float scale_p, float *__restrict__ out, int64_t ldo) {
    constexpr int RB = PER <= 8 ? 8 : 9;                         // bits of r
    constexpr uint32_t RMASK = (1u << RB) - 1u;
    extern __shared__ float smem_f[];
    float *t = smem_f;                                           // [n] descending activations of the neuron
    int32_t *rows = reinterpret_cast<int32_t *>(t + n);          // [n] their probe images
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *xs = t + 2 * n + warp * n;                            // [n] this warp's gathered column
    const int64_t j = blockIdx.y;
    const int c = blockIdx.x * 8 + warp;
    for (int r = threadIdx.x; r < n; r += 256) {
        t[r] = vals[int64_t(r) * K + j];
        rows[r] = idx[int64_t(r) * K + j];
    }
    __syncthreads();
    if (c >= C) return;
    uint32_t key[PER];
    float sum_x = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int r = lane * PER + e;
        key[e] = 0u;                                             // padding sorts last (no ordered key has zero top bits)
        if (r < n) {
            const float x = P[int64_t(rows[r]) * ldp + c];
            sum_x += x;
            xs[r] = x;
            key[e] = (ordered_key(x) & ~RMASK) | static_cast<uint32_t>(r);
        }
    }
    bitonic_desc_regs<PER>(key, lane);
    // neighbours with the same top bits?
    const uint32_t before = __shfl_up_sync(0xffffffffu, key[PER - 1], 1), after = __shfl_down_sync(0xffffffffu, key[0], 1);
    bool amb[PER];
    bool any_amb = false;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const uint32_t top = key[e] >> RB;
        const uint32_t prev = e > 0 ? key[e > 0 ? e - 1 : 0] : (lane > 0 ? before : ~key[e]);
        const uint32_t next = e + 1 < PER ? key[e + 1 < PER ? e + 1 : e] : (lane < 31 ? after : ~key[e]);
        amb[e] = key[e] != 0u && ((prev >> RB) == top || (next >> RB) == top);
        any_amb |= amb[e];
    }
    float acc = 0.f;
    if (!__any_sync(0xffffffffu, any_amb)) {
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int q = lane * PER + e;
            if (q < n) acc += abs_pow(t[key[e] & RMASK] - t[q], p);
        }
    } else {
        __syncwarp();                                            // xs[] written by other lanes
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            int q = lane * PER + e;
            unsigned pending = __ballot_sync(0xffffffffu, amb[e]);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint32_t w = __shfl_sync(0xffffffffu, key[e], src);
                const uint32_t r0 = w & RMASK, top0 = w >> RB, k0 = ordered_key(xs[r0]);
                int ahead = 0;                                   // elements before (k0, r0) in the descending (key, r) order
                for (int r = lane; r < n; r += 32) {
                    const uint32_t k = ordered_key(xs[r]);
                    ahead += ((k >> RB) > top0) || ((k >> RB) == top0 && (k > k0 || (k == k0 && uint32_t(r) > r0)));
                }
                ahead = __reduce_add_sync(0xffffffffu, ahead);
                if (lane == src) q = ahead;
            }
            if (lane * PER + e < n) acc += abs_pow(t[key[e] & RMASK] - t[q], p);
        }
    }
    sum_x = warp_sum(sum_x);
    acc = warp_sum(acc);
    if (lane == 0) {     // the neuron's baseline is applied by rank_normalize_kernel (it may still be in the making)
        const float avg = sum_x / static_cast<float>(n);
        out[j * ldo + c] = acc / static_cast<float>(n);
        den_out[j * C + c] = scale_p == 0.5f ? sqrtf(avg) : powf(avg, scale_p);
    }
}

// rank_sorted_cta_kernel: any n up to kRankMaxSort.  CTA = (neuron, group of `cpb` concepts); per concept the column's
// words are sorted in shared memory (bitonic network, one __syncthreads per stage).
constexpr int kRankMaxSort = 8192;

__global__ void __launch_bounds__(256)
rank_sorted_cta_kernel(const float *__restrict__ P, int64_t ldp, int C, const int32_t *__restrict__ idx,
                       const float *__restrict__ vals, int64_t K, int n, int npad, int cpb,
                       float *__restrict__ den_out, float p, float scale_p, float *__restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *key = reinterpret_cast<unsigned long long *>(smem_raw);     // [npad]
    float *t = reinterpret_cast<float *>(key + npad);                                // [n]
    int32_t *rows = reinterpret_cast<int32_t *>(t + n);                              // [n]
    __shared__ float red[2][8];
    const int64_t j = blockIdx.y;
    for (int r = threadIdx.x; r < n; r += 256) {
        t[r] = vals[int64_t(r) * K + j];
        rows[r] = idx[int64_t(r) * K + j];
    }
    __syncthreads();
    for (int cc = 0; cc < cpb; ++cc) {
        const int c = blockIdx.x * cpb + cc;
        if (c >= C) break;
        float sum_x = 0.f;
        for (int r = threadIdx.x; r < npad; r += 256) {
            unsigned long long w = 0ull;
            if (r < n) {
                const float x = P[int64_t(rows[r]) * ldp + c];
                sum_x += x;
                w = pack_key(ordered_key(x), static_cast<uint32_t>(r));
            }
            key[r] = w;
        }
        __syncthreads();
        for (int size = 2; size <= npad; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int tt = threadIdx.x; tt < (npad >> 1); tt += 256) {
                    const int i = 2 * tt - (tt & (stride - 1)), k2 = i + stride;
                    const bool desc = (i & size) == 0 || size == npad;
                    const unsigned long long a = key[i], b = key[k2];
                    if ((a < b) == desc) {
                        key[i] = b;
                        key[k2] = a;
                    }
                }
                __syncthreads();
            }
        float acc = 0.f;
        for (int q = threadIdx.x; q < n; q += 256) acc += abs_pow(t[static_cast<uint32_t>(key[q])] - t[q], p);
        sum_x = warp_sum(sum_x);
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = sum_x;
            red[1][threadIdx.x >> 5] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float sx = 0.f, se = 0.f;
            for (int w = 0; w < 8; ++w) {
                sx += red[0][w];
                se += red[1][w];
            }
            const float avg = sx / static_cast<float>(n);
            out[j * ldo + c] = se / static_cast<float>(n);
            den_out[j * C + c] = scale_p == 0.5f ? sqrtf(avg) : powf(avg, scale_p);
        }
        __syncthreads();
    }
}

// The reference's baseline: 5 x torch.randperm(n) per neuron from the global CPU generator.  torch.randperm is a
// Fisher-Yates shuffle that consumes one 32-bit Mersenne-Twister draw per step (z = draw % (n - i); swap i, i + z); the
// host replays the generator's raw draws (mammo_clip_dissect_b200/similarity.py) and this kernel runs the shuffles: one
// thread per (neuron, permutation), the working array in global scratch.  Position i is final after step i, so the
// term |asc[i] - asc[perm[i]]|^p is accumulated on the way.
__global__ void __launch_bounds__(128)
rank_perm_baseline_kernel(const float *__restrict__ vals, int64_t K, int n, const uint32_t *__restrict__ draws, float p,
                          int32_t *__restrict__ scratch, float *__restrict__ base_part) {
    const int64_t id = int64_t(blockIdx.x) * 128 + threadIdx.x;          // (neuron j, permutation s5)
    if (id >= K * 5) return;
    const int64_t j = id / 5;
    int32_t *perm = scratch + id * n;
    const uint32_t *d = draws + id * (n - 1);
    for (int i = 0; i < n; ++i) perm[i] = i;
    float acc = 0.f;
    auto asc = [&](int i) { return vals[int64_t(n - 1 - i) * K + j]; };
    for (int i = 0; i < n - 1; ++i) {
        const int z = i + static_cast<int>(d[i] % static_cast<uint32_t>(n - i));
        const int a = perm[i], b = perm[z];
        perm[z] = a;                                              // perm[i] = b is final: no need to store it
        acc += abs_pow(asc(i) - asc(b), p);
    }
    acc += abs_pow(asc(n - 1) - asc(perm[n - 1]), p);
    base_part[id] = acc;
}

// the same partial sums from host-supplied permutations (perms [K][5][n])
__global__ void __launch_bounds__(128)
rank_perm_given_kernel(const float *__restrict__ vals, int64_t K, int n, const int32_t *__restrict__ perms, float p,
                       float *__restrict__ base_part) {
    __shared__ float red[4];
    const int64_t id = blockIdx.x;                                        // (neuron j, permutation s5)
    const int64_t j = id / 5;
    float acc = 0.f;
    for (int r = threadIdx.x; r < n; r += 128) {
        const int q = perms[id * n + r];
        acc += abs_pow(vals[int64_t(n - 1 - r) * K + j] - vals[int64_t(n - 1 - q) * K + j], p);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) base_part[id] = red[0] + red[1] + red[2] + red[3];
}

// The generator itself on the device: MT19937 (the engine behind torch's CPU generator) continued from a given state.
// One CTA; a twist of the 624-word state runs as four data-parallel phases (word i needs the OLD words i, i+1 and the
// word i+397 mod 624, which is old for i < 227 and already new beyond), then up to 624 tempered outputs are written
// in parallel.  state_io: 624 words + the position of the next output (624 = "twist first"); left in place for the
// host to put back into the generator.  ~0.2 us per 624 draws: 640 k draws (c5) in 0.2 ms instead of 2.5 ms of numpy
// plus a host-to-device copy.
constexpr int kMtN = 624, kMtM = 397, kMtThreads = 256;

__global__ void __launch_bounds__(kMtThreads)
mt19937_draws_kernel(uint32_t *__restrict__ state_io, int64_t count, uint32_t *__restrict__ draws) {
    __shared__ uint32_t mt[kMtN];
    __shared__ int pos_s;
    for (int i = threadIdx.x; i < kMtN; i += kMtThreads) mt[i] = state_io[i];
    if (threadIdx.x == 0) pos_s = static_cast<int>(state_io[kMtN]);
    __syncthreads();
    int pos = pos_s;
    auto mix = [](uint32_t cur, uint32_t next, uint32_t far) {
        const uint32_t y = (cur & 0x80000000u) | (next & 0x7FFFFFFFu);
        return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    };
    int64_t done = 0;
    while (done < count) {
        if (pos >= kMtN) {
            // phases [0,227) [227,454) [454,623) {623}: read, barrier, write, barrier
            const int lo[4] = {0, kMtN - kMtM, 2 * (kMtN - kMtM), kMtN - 1}, hi[4] = {kMtN - kMtM, 2 * (kMtN - kMtM), kMtN - 1, kMtN};
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
                const int i = lo[ph] + threadIdx.x;
                uint32_t v = 0u;
                const bool on = i < hi[ph];
                if (on) v = mix(mt[i], mt[(i + 1) % kMtN], mt[(i + kMtM) % kMtN]);
                __syncthreads();
                if (on) mt[i] = v;
                __syncthreads();
            }
            pos = 0;
        }
        const int64_t left = count - done;
        const int take = static_cast<int>(left < int64_t(kMtN - pos) ? left : int64_t(kMtN - pos));
        for (int t = threadIdx.x; t < take; t += kMtThreads) {
            uint32_t y = mt[pos + t];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9D2C5680u;
            y ^= (y << 15) & 0xEFC60000u;
            y ^= y >> 18;
            draws[done + t] = y;
        }
        pos += take;
        done += take;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMtN; i += kMtThreads) state_io[i] = mt[i];
    if (threadIdx.x == 0) state_io[kMtN] = static_cast<uint32_t>(pos);
}

// out[j,c] = -((e / baseline_j) / den), e = mean_q |.|^p and den = (mean cosine)^scale_p left by the rank kernels: the
// reference's operation order (similarity.py:126-129), applied once the baseline kernel has delivered.
__global__ void __launch_bounds__(256)
rank_normalize_kernel(const float *__restrict__ base_part, const float *__restrict__ den, int C, int n, float *__restrict__ out,
                      int64_t ldo) {
    const int64_t j = blockIdx.y;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float base = 0.f;
    for (int s5 = 0; s5 < 5; ++s5) base += base_part[j * 5 + s5];
    base /= 5.f * static_cast<float>(n);
    const float err = out[j * ldo + c] / base;
    out[j * ldo + c] = -(err / den[j * C + c]);
}

static int launch_rank_sorted(const float *P, int64_t ldp, int64_t C, const int32_t *idx, const float *vals, int64_t K,
                              int64_t n, float *den, float p, float scale_p, float *out, int64_t ldo, cudaStream_t st) {
    if (n <= 512) {
        dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, 8)), static_cast<unsigned>(K));
        const size_t smem = size_t(n) * 8 + size_t(n) * 8 * 4;       // t, rows, and a column per warp
        if (n <= 256)
            rank_sorted_warp_kernel<8><<<grid, 256, smem, st>>>(P, ldp, int(C), idx, vals, K, int(n), den, p, scale_p, out, ldo);
        else
            rank_sorted_warp_kernel<16><<<grid, 256, smem, st>>>(P, ldp, int(C), idx, vals, K, int(n), den, p, scale_p, out, ldo);
        return check_launch();
    }
    int npad = 1024;
    while (npad < n) npad <<= 1;
    const size_t smem = size_t(npad) * 8 + size_t(n) * 8;
    if (cudaFuncSetAttribute(rank_sorted_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    // enough CTAs for the machine, as many concepts per CTA as that allows (t and rows are loaded once per CTA)
    int cpb = static_cast<int>((C * K) / (4 * int64_t(num_sms())));
    if (cpb < 1) cpb = 1;
    if (cpb > 16) cpb = 16;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, cpb)), static_cast<unsigned>(K));
    rank_sorted_cta_kernel<<<grid, 256, smem, st>>>(P, ldp, int(C), idx, vals, K, int(n), npad, cpb, den, p, scale_p, out, ldo);
    return check_launch();
}

static int launch_rank_normalize(const float *base_part, const float *den, int64_t K, int64_t C, int64_t n, float *out,
                                 int64_t ldo, cudaStream_t st) {
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, 256)), static_cast<unsigned>(K));
    rank_normalize_kernel<<<grid, 256, 0, st>>>(base_part, den, int(C), int(n), out, ldo);
    return check_launch();
}

}  // namespace mcd

namespace mcd {
struct RankWs {
    float *base_part;      // [5 K]  per (neuron, permutation) partial sums of the baseline
    float *den;            // [K C]  (mean cosine)^scale_p
    int32_t *scratch;      // [5 K top_n]  working arrays of the device-side shuffles
    size_t total;
};
static RankWs rank_ws(void *workspace, int64_t K, int64_t top_n, int64_t C) {
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    char *w = static_cast<char *>(workspace);
    RankWs r;
    const size_t o1 = up(size_t(K) * 5 * sizeof(float)), o2 = o1 + up(size_t(K) * size_t(C) * sizeof(float));
    r.base_part = reinterpret_cast<float *>(w);
    r.den = reinterpret_cast<float *>(w + o1);
    r.scratch = reinterpret_cast<int32_t *>(w + o2);
    r.total = o2 + size_t(K) * 5 * size_t(top_n) * sizeof(int32_t);
    return r;
}
static bool rank_args_ok(int64_t C, int64_t K, int64_t top_n) { return C >= 1 && K >= 1 && top_n >= 1; }
static bool rank_supported(int64_t C, int64_t K, int64_t top_n) { return top_n <= kRankMaxSort && K <= 65535 && C <= (1 << 24); }
}  // namespace mcd

extern "C" size_t mcd_rank_reorder_workspace_bytes(int64_t K, int64_t top_n, int64_t C) {
    if (K < 1 || top_n < 1 || C < 1) return 0;
    return mcd::rank_ws(nullptr, K, top_n, C).total;
}

extern "C" int mcd_mt19937_draws(uint32_t *state_io, int64_t count, uint32_t *draws, mcd_stream_t stream) {
    using namespace mcd;
    if (!state_io || !draws || count < 1) return MCD_ERR_INVALID_ARGUMENT;
    mt19937_draws_kernel<<<1, kMtThreads, 0, static_cast<cudaStream_t>(stream)>>>(state_io, count, draws);
    return check_launch();
}

extern "C" int mcd_rank_baseline_draws_f32(const float *vals, int64_t K, int64_t top_n, int64_t C, const uint32_t *draws,
                                           float p, void *workspace, size_t workspace_bytes, mcd_stream_t stream) {
    using namespace mcd;
    if (!vals || !draws || !workspace || !rank_args_ok(C, K, top_n)) return MCD_ERR_INVALID_ARGUMENT;
    if (!rank_supported(C, K, top_n)) return MCD_ERR_UNSUPPORTED;
    const RankWs w = rank_ws(workspace, K, top_n, C);
    if (workspace_bytes < w.total) return MCD_ERR_WORKSPACE;
    rank_perm_baseline_kernel<<<static_cast<unsigned>(ceil_div<int64_t>(K * 5, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        vals, K, int(top_n), draws, p, w.scratch, w.base_part);
    return check_launch();
}

extern "C" int mcd_rank_baseline_perms_f32(const float *vals, int64_t K, int64_t top_n, int64_t C, const int32_t *perms,
                                           float p, void *workspace, size_t workspace_bytes, mcd_stream_t stream) {
    using namespace mcd;
    if (!vals || !perms || !workspace || !rank_args_ok(C, K, top_n)) return MCD_ERR_INVALID_ARGUMENT;
    if (!rank_supported(C, K, top_n)) return MCD_ERR_UNSUPPORTED;
    const RankWs w = rank_ws(workspace, K, top_n, C);
    if (workspace_bytes < w.total) return MCD_ERR_WORKSPACE;
    rank_perm_given_kernel<<<static_cast<unsigned>(K * 5), 128, 0, static_cast<cudaStream_t>(stream)>>>(vals, K, int(top_n), perms, p,
                                                                                                        w.base_part);
    return check_launch();
}

extern "C" int mcd_rank_errors_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx, const float *vals,
                                   int64_t K, int64_t top_n, float p, float scale_p, void *workspace, size_t workspace_bytes,
                                   float *out, int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!P || !idx || !vals || !workspace || !out || N < 1 || !rank_args_ok(C, K, top_n) || ldp < C || ldo < C)
        return MCD_ERR_INVALID_ARGUMENT;
    if (!rank_supported(C, K, top_n)) return MCD_ERR_UNSUPPORTED;
    const RankWs w = rank_ws(workspace, K, top_n, C);
    if (workspace_bytes < w.total) return MCD_ERR_WORKSPACE;
    return launch_rank_sorted(P, ldp, C, idx, vals, K, top_n, w.den, p, scale_p, out, ldo, static_cast<cudaStream_t>(stream));
}

extern "C" int mcd_rank_finish_f32(int64_t K, int64_t C, int64_t top_n, void *workspace, size_t workspace_bytes, float *out,
                                   int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!workspace || !out || !rank_args_ok(C, K, top_n) || ldo < C) return MCD_ERR_INVALID_ARGUMENT;
    if (!rank_supported(C, K, top_n)) return MCD_ERR_UNSUPPORTED;
    const RankWs w = rank_ws(workspace, K, top_n, C);
    if (workspace_bytes < w.total) return MCD_ERR_WORKSPACE;
    return launch_rank_normalize(w.base_part, w.den, K, C, top_n, out, ldo, static_cast<cudaStream_t>(stream));
}

// the three steps on one stream
extern "C" int mcd_rank_reorder_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx,
                                    const float *vals, int64_t K, int64_t top_n, const int32_t *perms, float p,
                                    float scale_p, void *workspace, size_t workspace_bytes, float *out, int64_t ldo,
                                    mcd_stream_t stream) {
    int rc = mcd_rank_baseline_perms_f32(vals, K, top_n, C, perms, p, workspace, workspace_bytes, stream);
    if (rc == MCD_OK) rc = mcd_rank_errors_f32(P, ldp, N, C, idx, vals, K, top_n, p, scale_p, workspace, workspace_bytes, out, ldo, stream);
    if (rc == MCD_OK) rc = mcd_rank_finish_f32(K, C, top_n, workspace, workspace_bytes, out, ldo, stream);
    return rc;
}

extern "C" int mcd_rank_reorder_draws_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx,
                                          const float *vals, int64_t K, int64_t top_n, const uint32_t *draws, float p,
                                          float scale_p, void *workspace, size_t workspace_bytes, float *out, int64_t ldo,
                                          mcd_stream_t stream) {
    int rc = mcd_rank_baseline_draws_f32(vals, K, top_n, C, draws, p, workspace, workspace_bytes, stream);
    if (rc == MCD_OK) rc = mcd_rank_errors_f32(P, ldp, N, C, idx, vals, K, top_n, p, scale_p, workspace, workspace_bytes, out, ldo, stream);
    if (rc == MCD_OK) rc = mcd_rank_finish_f32(K, C, top_n, workspace, workspace_bytes, out, ldo, stream);
    return rc;
}
