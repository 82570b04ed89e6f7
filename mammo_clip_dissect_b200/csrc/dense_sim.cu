// Dense contractions of the scoring path, exact-fp32 CUDA-core form:
//   * I T^T with optional row normalisation     (concept_vit/utils.py:577-594)
//   * cos_similarity / cos_similarity_cubed      (concept_vit/similarity.py:7-47): column statistics
//     + A~^T P~ with the centring / cubing / scaling applied while the operand tile is loaded, so
//     the transformed copies of A and P that the reference materialises are never written.
// These kernels define the fp32 semantics (true fp32 multiply-accumulate like the reference's
// SGEMM with TF32 off); the tcgen05 path in gemm_tf32x3.cu is validated against them.
#include "common.cuh"

namespace mcd {

constexpr int kTile = 64, kTileK = 16, kGemmThreads = 256;

enum { kModeNT = 0, kModeCos = 1, kModeCos3 = 2 };

template <int MODE>
__device__ __forceinline__ float xform(float x, float s0, float s1) {
    if (MODE == kModeNT) return __fdiv_rn(x, s1);                      // x / ||row||   (s1 == 1: no normalisation)
    if (MODE == kModeCos) return __fdiv_rn(x, s1);                     // x / ||col||
    const float d = __fsub_rn(x, s0);                                   // ((x - mean)^3) / clip(||.||, min_norm)
    return __fdiv_rn(__fmul_rn(__fmul_rn(d, d), d), s1);
}

// C[m,n] = sum_k fa(A(k,m)) * fb(B(k,n)).
//   MODE NT : A(k,m) = A[m*lda + k], per-row scale a1[m]      (I T^T: A = I, B = T)
//   MODE Cos: A(k,m) = A[k*lda + m], per-column a0[m], a1[m]  (A^T P:  A = activations, B = P)
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads)
sgemm_xform_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ a0, const float *__restrict__ a1,
                   const float *__restrict__ B, int64_t ldb, const float *__restrict__ b0, const float *__restrict__ b1,
                   int64_t M, int64_t Nn, int64_t Kd, float *__restrict__ Cout, int64_t ldc) {
    __shared__ float As[kTileK][kTile + 4];
    __shared__ float Bs[kTileK][kTile + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = int64_t(blockIdx.y) * kTile, n0 = int64_t(blockIdx.x) * kTile;
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4] = {};

    for (int64_t k0 = 0; k0 < Kd; k0 += kTileK) {
        if (MODE == kModeNT) {
            const int mm = tid / 4, kk0 = (tid % 4) * 4;
            const int64_t m = m0 + mm, n = n0 + mm;
            const float sa = (m < M && a1) ? a1[m] : 1.f, sb = (n < Nn && b1) ? b1[n] : 1.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t kx = k0 + kk0 + e;
                As[kk0 + e][mm] = (m < M && kx < Kd) ? xform<MODE>(A[m * lda + kx], 0.f, sa) : 0.f;
                Bs[kk0 + e][mm] = (n < Nn && kx < Kd) ? xform<MODE>(B[n * ldb + kx], 0.f, sb) : 0.f;
            }
        } else {
            const int kk = tid / 16, mm0 = (tid % 16) * 4;
            const int64_t kx = k0 + kk;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t m = m0 + mm0 + e, n = n0 + mm0 + e;
                As[kk][mm0 + e] = (m < M && kx < Kd) ? xform<MODE>(A[kx * lda + m], a0 ? a0[m] : 0.f, a1[m]) : 0.f;
                Bs[kk][mm0 + e] = (n < Nn && kx < Kd) ? xform<MODE>(B[kx * ldb + n], b0 ? b0[n] : 0.f, b1[n]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTileK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[e] = As[kk][ty * 4 + e];
                b[e] = Bs[kk][tx * 4 + e];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n < Nn) Cout[m * ldc + n] = acc[i][j];
        }
    }
}

// ||row||_2 of X [R, D]: one warp per row
__global__ void __launch_bounds__(256) row_norm_kernel(const float *__restrict__ X, int64_t ldx, int64_t R, int64_t D,
                                                       float *__restrict__ norm) {
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (r >= R) return;
    float s = 0.f;
    for (int64_t d = lane; d < D; d += 32) {
        const float v = X[r * ldx + d];
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) norm[r] = sqrtf(s);
}

// column statistics of X [N, M]: thread per column, blocks of 32 columns x 8 row-lanes
__global__ void __launch_bounds__(256) col_stats_kernel(const float *__restrict__ X, int64_t ldx, int64_t N, int64_t M,
                                                        int cubed, float min_norm, float *__restrict__ mean_out,
                                                        float *__restrict__ norm_out) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t m = int64_t(blockIdx.x) * 32 + cx;
    float mean = 0.f;
    if (cubed) {
        float s = 0.f;
        if (m < M)
            for (int64_t i = ry; i < N; i += 8) s += X[i * ldx + m];
        red[ry][cx] = s;
        __syncthreads();
        if (ry == 0) {
            float t = 0.f;
            for (int q = 0; q < 8; ++q) t += red[q][cx];
            red[0][cx] = t / static_cast<float>(N);
        }
        __syncthreads();
        mean = red[0][cx];
        __syncthreads();
    }
    float s2 = 0.f;
    if (m < M)
        for (int64_t i = ry; i < N; i += 8) {
            float v = X[i * ldx + m];
            if (cubed) {
                const float d = __fsub_rn(v, mean);
                v = __fmul_rn(__fmul_rn(d, d), d);
            }
            s2 = fmaf(v, v, s2);
        }
    red[ry][cx] = s2;
    __syncthreads();
    if (ry == 0 && m < M) {
        float t = 0.f;
        for (int q = 0; q < 8; ++q) t += red[q][cx];
        float nrm = sqrtf(t);
        if (cubed) nrm = fmaxf(nrm, min_norm);           // torch.clip(norm, min_norm); NaN stays NaN
        if (cubed && t != t) nrm = t;
        norm_out[m] = nrm;
        if (mean_out) mean_out[m] = mean;
    }
}

}  // namespace mcd

extern "C" int mcd_col_stats_f32(const float *X, int64_t ldx, int64_t N, int64_t M, int cubed, float min_norm,
                                 float *mean_out, float *inv_norm_out, mcd_stream_t stream) {
    using namespace mcd;
    if (!X || !inv_norm_out || N < 1 || M < 1 || ldx < M || (cubed && !mean_out)) return MCD_ERR_INVALID_ARGUMENT;
    col_stats_kernel<<<static_cast<unsigned>(ceil_div<int64_t>(M, 32)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        X, ldx, N, M, cubed, min_norm, mean_out, inv_norm_out);
    return check_launch();
}

extern "C" int mcd_cos_matmul_f32(const float *A, int64_t lda, const float *meanA, const float *invA, const float *P,
                                  int64_t ldp, const float *meanP, const float *invP, int64_t N, int64_t K, int64_t C,
                                  int cubed, float *out, int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!A || !P || !invA || !invP || !out || N < 1 || K < 1 || C < 1 || lda < K || ldp < C || ldo < C)
        return MCD_ERR_INVALID_ARGUMENT;
    if (cubed && (!meanA || !meanP)) return MCD_ERR_INVALID_ARGUMENT;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, kTile)), static_cast<unsigned>(ceil_div<int64_t>(K, kTile)));
    if (grid.y > 65535) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cubed)
        sgemm_xform_kernel<kModeCos3><<<grid, kGemmThreads, 0, st>>>(A, lda, meanA, invA, P, ldp, meanP, invP, K, C, N,
                                                                    out, ldo);
    else
        sgemm_xform_kernel<kModeCos><<<grid, kGemmThreads, 0, st>>>(A, lda, nullptr, invA, P, ldp, nullptr, invP, K, C,
                                                                   N, out, ldo);
    return check_launch();
}

namespace mcd {
size_t cos_similarity_tc_workspace(int64_t N, int64_t K, int64_t C);                              // gemm_tf32x3.cu
int cos_similarity_tc(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K, int64_t C, int cubed,
                      float min_norm, float *out, int64_t ldo, void *ws, size_t ws_bytes, cudaStream_t st);
static int g_last_cos_path = 0;
}  // namespace mcd

extern "C" int mcd_last_cos_path(void) { return mcd::g_last_cos_path; }

extern "C" size_t mcd_cos_similarity_workspace_bytes(int64_t N, int64_t K, int64_t C) {
    if (N < 1 || K < 1 || C < 1) return 0;
    // the tensor-core path's buffers, or the CUDA-core path's column statistics (2 x (K + C) floats)
    const size_t tc = mcd::cos_similarity_tc_workspace(N, K, C), cc = size_t(2) * size_t(K + C) * sizeof(float) + 256;
    return tc > cc ? tc : cc;
}

// The whole cos_similarity / cos_similarity_cubed call (similarity.py:7-47) behind one entry point: column statistics of
// both matrices, then out [K, C] = f(A)^T f(P) on the tensor cores (gemm_tf32x3.cu).  Falls back to the exact CUDA-core
// kernels only when the tensor-map encoder is unavailable or tunable gemm_variant = 1 asks for it (mcd_last_cos_path tells).
extern "C" int mcd_cos_similarity_f32(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K,
                                      int64_t C, int cubed, float min_norm, float *out, int64_t ldo, void *workspace,
                                      size_t workspace_bytes, mcd_stream_t stream) {
    using namespace mcd;
    if (!P || !A || !out || N < 1 || K < 1 || C < 1 || ldp < C || lda < K || ldo < C) return MCD_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < mcd_cos_similarity_workspace_bytes(N, K, C)) return MCD_ERR_WORKSPACE;
    int rc = MCD_ERR_UNSUPPORTED;
    if (tunable(kGemmVariant) != 1) {
        rc = cos_similarity_tc(P, ldp, A, lda, N, K, C, cubed, min_norm, out, ldo, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
        if (rc == MCD_OK) g_last_cos_path = 1;
    }
    if (rc == MCD_ERR_UNSUPPORTED) {
        float *st = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
        float *meanP = st, *normP = st + C, *meanA = st + 2 * C, *normA = st + 2 * C + K;
        rc = mcd_col_stats_f32(P, ldp, N, C, cubed, min_norm, meanP, normP, stream);
        if (rc != MCD_OK) return rc;
        rc = mcd_col_stats_f32(A, lda, N, K, cubed, min_norm, meanA, normA, stream);
        if (rc != MCD_OK) return rc;
        rc = mcd_cos_matmul_f32(A, lda, meanA, normA, P, ldp, meanP, normP, N, K, C, cubed, out, ldo, stream);
        if (rc == MCD_OK) g_last_cos_path = 3;
    }
    return rc;
}

// fp32 CUDA-core form of K1 (exact-fp32 reference semantics).  workspace: (N + C) floats of norms.
namespace mcd {
int sim_matrix_fp32(const float *I, int64_t ldi, const float *T, int64_t ldt, int64_t N, int64_t C, int64_t D,
                    int normalize_rows, float *P, int64_t ldp, float *norms, cudaStream_t st) {
    const float *nI = nullptr, *nT = nullptr;
    if (normalize_rows) {
        row_norm_kernel<<<static_cast<unsigned>(ceil_div<int64_t>(N, 8)), 256, 0, st>>>(I, ldi, N, D, norms);
        row_norm_kernel<<<static_cast<unsigned>(ceil_div<int64_t>(C, 8)), 256, 0, st>>>(T, ldt, C, D, norms + N);
        count_launch(2);
        nI = norms;
        nT = norms + N;
    }
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, kTile)), static_cast<unsigned>(ceil_div<int64_t>(N, kTile)));
    if (grid.y > 65535) return MCD_ERR_UNSUPPORTED;
    sgemm_xform_kernel<kModeNT><<<grid, kGemmThreads, 0, st>>>(I, ldi, nullptr, nI, T, ldt, nullptr, nT, N, C, D, P, ldp);
    return check_launch();
}
}  // namespace mcd

namespace mcd {
size_t sim_matrix_tc_workspace(int64_t N, int64_t C, int64_t D);                       // gemm_tf32x3.cu
int sim_matrix_tc(const float *I, int64_t ldi, const float *T, int64_t ldt, int64_t N, int64_t C, int64_t D,
                  int normalize_rows, float *P, int64_t ldp, float *S, int64_t lds, float a, void *ws, size_t ws_bytes,
                  int kind, int *S_done, cudaStream_t st);
static int g_last_gemm_path = 0;        // what the last mcd_gemm_nt_softmax_f32 call ran (mcd_last_gemm_path)
static size_t gemm_base_workspace(int64_t N, int64_t C) {
    const size_t ldp = size_t(ceil_div<int64_t>(C, 4) * 4);
    return ((size_t(N) + size_t(C)) * sizeof(float) + 255) / 256 * 256 + size_t(N) * ldp * sizeof(float) + 256;
}
}  // namespace mcd

extern "C" int mcd_last_gemm_path(void) { return mcd::g_last_gemm_path; }

extern "C" size_t mcd_gemm_nt_softmax_workspace_bytes(int64_t N, int64_t C, int64_t D) {
    if (N < 1 || C < 1 || D < 1) return 0;
    return mcd::gemm_base_workspace(N, C) + mcd::sim_matrix_tc_workspace(N, C, D);
}

extern "C" int mcd_gemm_nt_softmax_f32(const float *I, int64_t ldi, const float *T, int64_t ldt, int64_t N, int64_t C,
                                       int64_t D, int normalize_rows, float a, float *P_out, int64_t ldp, float *S_out,
                                       int64_t lds, void *workspace, size_t workspace_bytes, mcd_stream_t stream) {
    using namespace mcd;
    if (!I || !T || N < 1 || C < 1 || D < 1 || ldi < D || ldt < D || (!P_out && !S_out)) return MCD_ERR_INVALID_ARGUMENT;
    if ((P_out && ldp < C) || (S_out && lds < C)) return MCD_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < mcd_gemm_nt_softmax_workspace_bytes(N, C, D)) return MCD_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *norms = static_cast<float *>(workspace);
    float *P = P_out;
    int64_t ld = ldp;
    if (!P) {
        const size_t off = ((size_t(N) + size_t(C)) * sizeof(float) + 255) / 256 * 256;
        P = reinterpret_cast<float *>(static_cast<char *>(workspace) + off);
        ld = ceil_div<int64_t>(C, 4) * 4;
    }
    // tensor-core path (tcgen05 kind::tf32, 3-term split).  Default: the streaming kernel (a CTA walks all column tiles of
    // its 128-row band; the epilogue of a tile runs under the MMAs of the next) followed by the stand-alone softmax.
    // Tunable gemm_variant: 1 = the fp32 CUDA-core kernel; 2 = one CTA per output tile (round 1); 3 = the band kernel that
    // also rescales the band inside the GEMM kernel; 4 = the streaming kernel whose epilogue keeps every row's online
    // softmax pair + a normalising pass.  Measured at N = 100k (tools/bench_k1.py): 0.77 + 0.10 ms (default), the same
    // (2), 3.05 ms (3), 1.03 ms (4): the fused forms are correct and tested but lose -- the exp work lands on the four
    // epilogue warps of an SM whose other warp slots are empty, while the stand-alone kernel uses the whole machine.
    int rc = MCD_ERR_UNSUPPORTED;
    const int64_t variant = tunable(kGemmVariant);
    if (variant != 1) {
        char *tc_ws = static_cast<char *>(workspace) + gemm_base_workspace(N, C);
        int S_done = 0;
        rc = sim_matrix_tc(I, ldi, T, ldt, N, C, D, normalize_rows, P, ld, S_out, lds, a, tc_ws,
                           workspace_bytes - gemm_base_workspace(N, C), variant >= 2 && variant <= 4 ? int(variant) : 0, &S_done, st);
        if (rc == MCD_OK) g_last_gemm_path = S_done ? 2 : 1;
        if (rc == MCD_OK && S_done) return rc;
    }
    if (rc == MCD_ERR_UNSUPPORTED) {
        rc = sim_matrix_fp32(I, ldi, T, ldt, N, C, D, normalize_rows, P, ld, norms, st);
        if (rc == MCD_OK) g_last_gemm_path = 3;
    }
    if (rc != MCD_OK || !S_out) return rc;
    return mcd_softmax_rows_f32(P, ld, S_out, lds, N, C, a, stream);
}
