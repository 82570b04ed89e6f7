// Per-neuron top concepts: the t largest entries of every row of the score matrix, sorted descending
// (replaces torch.topk(similarities, k=10, dim=1) / torch.max(similarities, dim=1) of the reference's callers,
//  concept_vit/describe_broad_neurons.py:101, describe_clip_neurons.py:64).
//
// Total order as for the column top-k: value descending, concept index ascending, NaN largest, -0.0 == +0.0
// (torch.topk leaves the order of equal values unspecified).  One warp per row: a lane keeps every 32nd entry of the
// row in registers as an ordered key, and t rounds of a warp arg-max (64-bit (key, ~index) words, butterfly
// shuffles) emit the winners in order; the winner's lane retires its entry.
#include "common.cuh"

namespace mcd {

constexpr int kRowTopkWarps = 8;

template <int PER>     // rows up to 32 * PER entries
__global__ void __launch_bounds__(kRowTopkWarps * 32)
row_topk_kernel(const float *__restrict__ X, int64_t ldx, int64_t n_rows, int n_cols, int t,
                float *__restrict__ vals, int64_t *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * kRowTopkWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float *src = X + row * ldx;
    unsigned long long w[PER];          // (ordered key, ~column); 0 = no entry
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        w[i] = c < n_cols ? pack_key(ordered_key(__ldg(src + c)), ~static_cast<uint32_t>(c)) : 0ull;
    }
    for (int r = 0; r < t; ++r) {
        unsigned long long best = w[0];
#pragma unroll
        for (int i = 1; i < PER; ++i) best = w[i] > best ? w[i] : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        const uint32_t col = ~static_cast<uint32_t>(best);
        if (lane == 0) {
            idx[row * t + r] = static_cast<int64_t>(col);
            vals[row * t + r] = src[col];                 // the input bits (NaN payloads, signed zeros)
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) w[i] = w[i] == best ? 0ull : w[i];     // entries are distinct: exactly one retires
    }
}

}  // namespace mcd

extern "C" int mcd_row_topk_f32(const float *X, int64_t ldx, int64_t n_rows, int64_t n_cols, int64_t t, float *vals_out,
                                int64_t *idx_out, mcd_stream_t stream) {
    using namespace mcd;
    if (!X || !vals_out || !idx_out || n_rows < 1 || n_cols < 1 || ldx < n_cols || t < 1 || t > n_cols)
        return MCD_ERR_INVALID_ARGUMENT;
    if (n_cols > 32 * 32 || t > 64) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_rows, kRowTopkWarps));
    const int threads = kRowTopkWarps * 32, nc = static_cast<int>(n_cols), ti = static_cast<int>(t);
    if (n_cols <= 32 * 8)
        row_topk_kernel<8><<<grid, threads, 0, st>>>(X, ldx, n_rows, nc, ti, vals_out, idx_out);
    else if (n_cols <= 32 * 24)
        row_topk_kernel<24><<<grid, threads, 0, st>>>(X, ldx, n_rows, nc, ti, vals_out, idx_out);
    else
        row_topk_kernel<32><<<grid, threads, 0, st>>>(X, ldx, n_rows, nc, ti, vals_out, idx_out);
    return check_launch();
}
