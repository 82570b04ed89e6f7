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
// rank_reorder_kernel: one CTA per (neuron, 32-concept tile).  The [top_n x 32] slab of cosines sits in shared
// memory; thread (warp w, lane c) ranks rows w, w+8, ... of concept c by counting (top_n compares each, the 32
// lanes of a warp read one slab row per step: conflict-free).
#include "common.cuh"

namespace mcd {

constexpr int kRankThreads = 256, kRankTile = 32, kRankMaxTop = 512;

__device__ __forceinline__ float abs_pow(float d, float p) {
    const float a = fabsf(d);
    if (p == 3.f) return a * a * a;
    if (p == 2.f) return a * a;
    if (p == 1.f) return a;
    return powf(a, p);
}

__global__ void __launch_bounds__(128)
rank_baseline_kernel(const float *__restrict__ vals, int64_t K, int top_n, const int32_t *__restrict__ perms, float p,
                     float *__restrict__ baseline) {
    __shared__ float red[4];
    const int64_t j = blockIdx.x;
    const int total = 5 * top_n;
    float acc = 0.f;
    for (int i = threadIdx.x; i < total; i += 128) {
        const int s = i / top_n, r = i - s * top_n;
        const int q = perms[(j * 5 + s) * top_n + r];
        const float a = vals[int64_t(top_n - 1 - r) * K + j], b = vals[int64_t(top_n - 1 - q) * K + j];
        acc += abs_pow(a - b, p);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) baseline[j] = (red[0] + red[1] + red[2] + red[3]) / static_cast<float>(total);
}

__global__ void __launch_bounds__(kRankThreads)
rank_reorder_kernel(const float *__restrict__ P, int64_t ldp, int C, const int32_t *__restrict__ idx,
                    const float *__restrict__ vals, int64_t K, int top_n, const float *__restrict__ baseline, float p,
                    float scale_p, float *__restrict__ out, int64_t ldo) {
    extern __shared__ float smem_f[];
    float *x = smem_f;                                 // [top_n][32]
    float *t = x + top_n * kRankTile;                  // [top_n] descending activations
    float *red = t + top_n;                            // [2][8][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = blockIdx.y;
    const int c = blockIdx.x * kRankTile + lane;
    for (int r = threadIdx.x; r < top_n; r += kRankThreads) t[r] = vals[int64_t(r) * K + j];
    for (int r = warp; r < top_n; r += kRankThreads / 32)
        x[r * kRankTile + lane] = c < C ? P[int64_t(idx[int64_t(r) * K + j]) * ldp + c] : 0.f;
    __syncthreads();
    float sum_x = 0.f, acc = 0.f;
    for (int r = warp; r < top_n; r += kRankThreads / 32) {
        const float mine = x[r * kRankTile + lane];
        sum_x += mine;
        int rank = 0;
        for (int q = 0; q < top_n; ++q) {
            const float v = x[q * kRankTile + lane];
            rank += (v < mine) || (v == mine && q < r);
        }
        acc += abs_pow(t[r] - t[top_n - 1 - rank], p);
    }
    red[warp * 32 + lane] = sum_x;
    red[256 + warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0 && c < C) {
        float sx = 0.f, se = 0.f;
#pragma unroll
        for (int w = 0; w < kRankThreads / 32; ++w) {
            sx += red[w * 32 + lane];
            se += red[256 + w * 32 + lane];
        }
        const float avg = sx / static_cast<float>(top_n);
        const float err = (se / static_cast<float>(top_n)) / baseline[j];
        const float den = scale_p == 0.5f ? sqrtf(avg) : powf(avg, scale_p);
        out[j * ldo + c] = -(err / den);
    }
}

}  // namespace mcd

extern "C" int mcd_rank_reorder_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx,
                                    const float *vals, int64_t K, int64_t top_n, const int32_t *perms, float p,
                                    float scale_p, float *baseline_ws, float *out, int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!P || !idx || !vals || !perms || !baseline_ws || !out || N < 1 || C < 1 || K < 1 || top_n < 1 || ldp < C || ldo < C)
        return MCD_ERR_INVALID_ARGUMENT;
    if (top_n > kRankMaxTop || K > 65535 || C > (1 << 24)) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rank_baseline_kernel<<<static_cast<unsigned>(K), 128, 0, st>>>(vals, K, int(top_n), perms, p, baseline_ws);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    const size_t smem = (size_t(top_n) * kRankTile + size_t(top_n) + 512) * sizeof(float);
    if (cudaFuncSetAttribute(rank_reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
        return MCD_ERR_CUDA;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, kRankTile)), static_cast<unsigned>(K));
    rank_reorder_kernel<<<grid, kRankThreads, smem, st>>>(P, ldp, int(C), idx, vals, K, int(top_n), baseline_ws, p, scale_p,
                                                         out, ldo);
    return check_launch();
}
