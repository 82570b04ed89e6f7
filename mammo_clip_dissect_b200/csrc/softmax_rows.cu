// K1b -- S = softmax(a * P, dim=1) for the concept-probability matrix
// (replaces torch.nn.functional.softmax(a*clip_feats, dim=1), concept_vit/similarity.py:54 / :80).
//
// One warp per probe image.  The row (C = 763 floats) is read once into registers, the scaled
// logits use a separately rounded multiply (a*x as the reference computes it), exp is the
// full-precision expf and the normalisation is a true division, so each element goes through the
// reference's operation sequence; only the order of the row sum (warp tree) differs.
// S is written with its own leading dimension (the product pads rows to a multiple of 4 floats
// so K3 can gather with aligned 16-byte loads); padding columns are zero-filled.
#include "common.cuh"

namespace mcd {

constexpr int kSoftmaxWarps = 8;

template <int PER, int kFull>   // register-resident rows up to 32*PER columns; the first 32*kFull exist in every row
__global__ void __launch_bounds__(kSoftmaxWarps * 32)
softmax_rows_kernel(const float *__restrict__ P, int64_t ldp, float *__restrict__ S, int64_t lds, int64_t n_rows,
                    int n_cols, float a) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * kSoftmaxWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float *src = P + row * ldp;
    float *dst = S + row * lds;
    float x[PER];
    float m = -INFINITY;
    // columns below 32 * kFull exist in every row (the dispatcher guarantees n_cols > 32 * kFull): only the last
    // register slots carry a bounds check
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        x[i] = (i < kFull || c < n_cols) ? __fmul_rn(a, __ldg(src + c)) : -INFINITY;
        m = fmaxf(m, x[i]);
    }
    m = warp_max(m);
    float sum = 0.f, low = 1.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        const bool have = i < kFull || c < n_cols;
        x[i] = have ? expf(__fsub_rn(x[i], m)) : 0.f;
        low = fminf(low, have ? x[i] : 1.f);
        sum += x[i];
    }
    sum = warp_sum(sum);
    // x / sum, correctly rounded, from the shared reciprocal: q0 = x * r, rem = x - q0 * sum (exact in one FMA),
    // q = q0 + rem * r (Markstein).  It equals the IEEE quotient whenever r = RN(1 / sum) is good enough, which fails only
    // for divisors with an all-ones mantissa and for quotients near the denormal range: such rows (a warp-uniform
    // decision) take the division instruction sequence.
    const float r = __frcp_rn(sum);
    const bool plain_div = __any_sync(0xFFFFFFFFu, !(low >= 1e-30f)) || (__float_as_uint(sum) & 0x7FFFFFu) == 0x7FFFFFu ||
                           !(sum < 3.0e38f) || !(m > -3.0e38f);
    if (plain_div) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            if (i < kFull || c < n_cols) dst[c] = __fdiv_rn(x[i], sum);
            else if (c < lds) dst[c] = 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            const float q0 = __fmul_rn(x[i], r);
            if (i < kFull || c < n_cols) dst[c] = __fmaf_rn(__fmaf_rn(-q0, sum, x[i]), r, q0);
            else if (c < lds) dst[c] = 0.f;
        }
    }
    for (int c = 32 * PER + lane; c < lds; c += 32) dst[c] = 0.f;
}

// any width: three passes over the row (it stays in L1/L2)
__global__ void __launch_bounds__(kSoftmaxWarps * 32)
softmax_rows_wide_kernel(const float *__restrict__ P, int64_t ldp, float *__restrict__ S, int64_t lds,
                         int64_t n_rows, int64_t n_cols, float a) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * kSoftmaxWarps + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float *src = P + row * ldp;
    float *dst = S + row * lds;
    float m = -INFINITY;
    for (int64_t c = lane; c < n_cols; c += 32) m = fmaxf(m, __fmul_rn(a, src[c]));
    m = warp_max(m);
    float sum = 0.f;
    for (int64_t c = lane; c < n_cols; c += 32) {
        const float e = expf(__fsub_rn(__fmul_rn(a, src[c]), m));
        dst[c] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    for (int64_t c = lane; c < n_cols; c += 32) dst[c] = __fdiv_rn(dst[c], sum);
    for (int64_t c = n_cols + lane; c < lds; c += 32) dst[c] = 0.f;
}

}  // namespace mcd

extern "C" int mcd_softmax_rows_f32(const float *P, int64_t ldp, float *S, int64_t lds, int64_t n_rows,
                                    int64_t n_cols, float a, mcd_stream_t stream) {
    using namespace mcd;
    if (!P || !S || n_rows < 1 || n_cols < 1 || ldp < n_cols || lds < n_cols) return MCD_ERR_INVALID_ARGUMENT;
    if (P == S) return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_rows, kSoftmaxWarps));
    const int threads = kSoftmaxWarps * 32;
    const int nc = static_cast<int>(n_cols);
    if (n_cols <= 32 * 4)
        launch_pdl((softmax_rows_kernel<4, 0>), dim3(grid), dim3(threads), 0, st, P, ldp, S, lds, n_rows, nc, a);
    else if (n_cols <= 32 * 22)
        launch_pdl((softmax_rows_kernel<24, 0>), dim3(grid), dim3(threads), 0, st, P, ldp, S, lds, n_rows, nc, a);
    else if (n_cols <= 32 * 24)          // the 763-concept set: 22 full register slots, bounds checks on the last two
        launch_pdl((softmax_rows_kernel<24, 22>), dim3(grid), dim3(threads), 0, st, P, ldp, S, lds, n_rows, nc, a);
    else if (n_cols <= 32 * 64)
        launch_pdl((softmax_rows_kernel<64, 0>), dim3(grid), dim3(threads), 0, st, P, ldp, S, lds, n_rows, nc, a);
    else
        launch_pdl((softmax_rows_wide_kernel), dim3(grid), dim3(threads), 0, st, P, ldp, S, lds, n_rows, n_cols, a);
    return check_launch();
}
