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
template <int PER>
__global__ void __launch_bounds__(256)
rank_sorted_warp_kernel(const float *__restrict__ P, int64_t ldp, int C, const int32_t *__restrict__ idx,
                        const float *__restrict__ vals, int64_t K, int n, const float *__restrict__ base_part, float p,
                        float scale_p, float *__restrict__ out, int64_t ldo) {
    extern __shared__ float smem_f[];
    float *t = smem_f;                                           // [n] descending activations of the neuron
    int32_t *rows = reinterpret_cast<int32_t *>(t + n);          // [n] their probe images
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = blockIdx.y;
    const int c = blockIdx.x * 8 + warp;
    for (int r = threadIdx.x; r < n; r += 256) {
        t[r] = vals[int64_t(r) * K + j];
        rows[r] = idx[int64_t(r) * K + j];
    }
    __syncthreads();
    if (c >= C) return;
    unsigned long long key[PER];
    float sum_x = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int r = lane * PER + e;
        key[e] = 0ull;                                           // padding sorts last
        if (r < n) {
            const float x = P[int64_t(rows[r]) * ldp + c];
            sum_x += x;
            key[e] = pack_key(ordered_key(x), static_cast<uint32_t>(r));
        }
    }
    bitonic_desc_regs<PER>(key, lane);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int q = lane * PER + e;
        if (q < n) acc += abs_pow(t[static_cast<uint32_t>(key[e])] - t[q], p);
    }
    sum_x = warp_sum(sum_x);
    acc = warp_sum(acc);
    if (lane == 0) {
        float base = 0.f;
        for (int s5 = 0; s5 < 5; ++s5) base += base_part[j * 5 + s5];
        base /= 5.f * static_cast<float>(n);
        const float avg = sum_x / static_cast<float>(n);
        const float err = (acc / static_cast<float>(n)) / base;
        const float den = scale_p == 0.5f ? sqrtf(avg) : powf(avg, scale_p);
        out[j * ldo + c] = -(err / den);
    }
}

// rank_sorted_cta_kernel: any n up to kRankMaxSort.  CTA = (neuron, group of `cpb` concepts); per concept the column's
// words are sorted in shared memory (bitonic network, one __syncthreads per stage).
constexpr int kRankMaxSort = 8192;

__global__ void __launch_bounds__(256)
rank_sorted_cta_kernel(const float *__restrict__ P, int64_t ldp, int C, const int32_t *__restrict__ idx,
                       const float *__restrict__ vals, int64_t K, int n, int npad, int cpb,
                       const float *__restrict__ base_part, float p, float scale_p, float *__restrict__ out, int64_t ldo) {
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
    float base = 0.f;
    for (int s5 = 0; s5 < 5; ++s5) base += base_part[j * 5 + s5];
    base /= 5.f * static_cast<float>(n);
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
            const float err = (se / static_cast<float>(n)) / base;
            const float den = scale_p == 0.5f ? sqrtf(avg) : powf(avg, scale_p);
            out[j * ldo + c] = -(err / den);
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

static int launch_rank_sorted(const float *P, int64_t ldp, int64_t C, const int32_t *idx, const float *vals, int64_t K,
                              int64_t n, const float *base_part, float p, float scale_p, float *out, int64_t ldo,
                              cudaStream_t st) {
    if (n <= 512) {
        dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, 8)), static_cast<unsigned>(K));
        const size_t smem = size_t(n) * 8;
        if (n <= 256)
            rank_sorted_warp_kernel<8><<<grid, 256, smem, st>>>(P, ldp, int(C), idx, vals, K, int(n), base_part, p, scale_p, out, ldo);
        else
            rank_sorted_warp_kernel<16><<<grid, 256, smem, st>>>(P, ldp, int(C), idx, vals, K, int(n), base_part, p, scale_p, out, ldo);
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
    rank_sorted_cta_kernel<<<grid, 256, smem, st>>>(P, ldp, int(C), idx, vals, K, int(n), npad, cpb, base_part, p, scale_p, out, ldo);
    return check_launch();
}

}  // namespace mcd

extern "C" int mcd_rank_reorder_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx,
                                    const float *vals, int64_t K, int64_t top_n, const int32_t *perms, float p,
                                    float scale_p, float *baseline_ws, float *out, int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!P || !idx || !vals || !perms || !baseline_ws || !out || N < 1 || C < 1 || K < 1 || top_n < 1 || ldp < C || ldo < C)
        return MCD_ERR_INVALID_ARGUMENT;
    if (top_n > kRankMaxSort || K > 65535 || C > (1 << 24)) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rank_perm_given_kernel<<<static_cast<unsigned>(K * 5), 128, 0, st>>>(vals, K, int(top_n), perms, p, baseline_ws);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    return launch_rank_sorted(P, ldp, C, idx, vals, K, top_n, baseline_ws, p, scale_p, out, ldo, st);
}

extern "C" size_t mcd_rank_reorder_workspace_bytes(int64_t K, int64_t top_n) {
    if (K < 1 || top_n < 1) return 0;
    return size_t(K) * 5 * sizeof(float) + 256 + size_t(K) * 5 * size_t(top_n) * sizeof(int32_t);
}

extern "C" int mcd_rank_reorder_draws_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx,
                                          const float *vals, int64_t K, int64_t top_n, const uint32_t *draws, float p,
                                          float scale_p, void *workspace, size_t workspace_bytes, float *out, int64_t ldo,
                                          mcd_stream_t stream) {
    using namespace mcd;
    if (!P || !idx || !vals || !draws || !workspace || !out || N < 1 || C < 1 || K < 1 || top_n < 1 || ldp < C || ldo < C)
        return MCD_ERR_INVALID_ARGUMENT;
    if (top_n > kRankMaxSort || K > 65535 || C > (1 << 24)) return MCD_ERR_UNSUPPORTED;
    if (workspace_bytes < mcd_rank_reorder_workspace_bytes(K, top_n)) return MCD_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *base_part = static_cast<float *>(workspace);
    int32_t *scratch = reinterpret_cast<int32_t *>(static_cast<char *>(workspace) + (size_t(K) * 5 * sizeof(float) + 255) / 256 * 256);
    rank_perm_baseline_kernel<<<static_cast<unsigned>(ceil_div<int64_t>(K * 5, 128)), 128, 0, st>>>(vals, K, int(top_n), draws, p,
                                                                                                    scratch, base_part);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    return launch_rank_sorted(P, ldp, C, idx, vals, K, top_n, base_part, p, scale_p, out, ldo, st);
}
