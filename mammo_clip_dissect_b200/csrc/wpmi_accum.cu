// K3 -- fused gather + rank-weighted log-sum: the body of soft-WPMI / WPMI
// (replaces the per-neuron Python loop of concept_vit/similarity.py:59-65 and :85-89).
//
//   L[j,c] = sum_r log( 1 + p[r] * (S[idx[r,j], c] - 1) + min_prob )        (soft-WPMI)
//   L[j,c] = sum_r log( S[idx[r,j], c] + min_prob )                         (WPMI, p == NULL)
//
// The log is MUFU lg2 accumulated in the log2 domain and scaled by ln2 once per output; by default the terms of four
// consecutive ranks are multiplied before one lg2, and a term is evaluated as one FMA (see term<> below; both are
// ~1e-7 relative effects on L against a stated tolerance of 1e-5).  The reference's own per-element operation order
// (sub, mul, add, add -- no FMA contraction) is kept selectable and is what runs whenever the grouping
// preconditions do not hold.
//
// Work decomposition: a CTA of 192 threads handles NPB = 192/TPN neurons for one tile of
// 4*TPN concepts; each thread owns 4 adjacent concepts (one 16-byte load per gathered row) and
// walks the k gathered rows with kU independent loads in flight.  Concept tiles are the slowest
// grid dimension so that, when a tile's slice of S fits in L2, the gathers of all neurons hit it
// before the next slice is touched (S is 305 MB at N=100k: larger than the 126 MB L2).
#include "common.cuh"
#include "topk_api.cuh"

namespace mcd {

constexpr int kAccumThreads = 192;
constexpr int kAccumMaxK = 512;

// lg2.approx without .ftz makes the compiler wrap every MUFU in a denormal rescue (compare, scale by 2^24,
// subtract 24): three extra instructions per element in an issue-bound kernel.  With min_prob a normal number the
// argument (>= min_prob for probabilities in [0,1] and weights <= 1) can never be subnormal, so the host selects FTZ.
template <bool FTZ>
__device__ __forceinline__ float lg2_fast(float x) {
    float r;
    if (FTZ) asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    else asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// value inside the log.  FUSED = false: the reference's unfused sequence (sub, mul, add, add), bit for bit.
// FUSED = true: the same quantity as one FMA, p*S + c with c = fl32(1 - p + eps) prepared in double per rank -- it is
// closer to the exact value than the reference's own fp32 sequence (whose rounding of S - 1 costs up to 3e-8 absolute
// on terms as small as 2e-3) and differs from it by ~2e-5 absolute in a sum of 100 logs of magnitude ~450
// (5e-8 relative; the stated tolerance is 1e-5; lg2.approx alone contributes 1.4e-5).
template <bool SOFT, bool FUSED>
__device__ __forceinline__ float term(float s, float w, float c, float eps) {
    if (SOFT) {
        if (FUSED) return __fmaf_rn(w, s, c);
        float v = __fsub_rn(s, 1.0f);
        v = __fmul_rn(w, v);
        v = __fadd_rn(1.0f, v);
        return __fadd_rn(v, eps);
    }
    return __fadd_rn(s, eps);
}

template <bool VEC>
__device__ __forceinline__ float4 load_row(const char *base, uint32_t off, int nvalid, uint64_t keep) {
    const float *row = reinterpret_cast<const float *>(base + off);
    if (VEC) return ldg_nc_v4_hint(row, keep);
    float4 s;
    s.x = __ldg(row);
    s.y = nvalid > 1 ? __ldg(row + 1) : 0.f;
    s.z = nvalid > 2 ? __ldg(row + 2) : 0.f;
    s.w = nvalid > 3 ? __ldg(row + 3) : 0.f;
    return s;
}

// The MUFU pipe (16 lg2 / clk / SM) is the binding unit of this kernel when every term gets its own lg2
// (k*K*C = 2.5e9 logs at c4 = 0.54 ms).  With GROUPED the terms of 4 consecutive ranks are multiplied first:
// log(t0 t1 t2 t3), one lg2 per 4 terms.  Valid when every term lies in [eps, 1 + eps] with eps >= 1e-9 (no
// underflow: the product is >= 1e-36; no sign cancellation), i.e. for probabilities S in [0,1] and weights p in
// [0,1]; the host checks eps, the CTA checks p, S in [0,1] is the documented domain of this entry point.  The extra
// rounding (3 products, 6e-8 relative each = 1.8e-7 absolute in the log) is below lg2.approx's own error.
template <int TPN, int U, bool SOFT, bool VEC, bool FTZ, bool GROUPED, bool FUSED>
__global__ void __launch_bounds__(kAccumThreads)
wpmi_accum_kernel(const float *__restrict__ S, int64_t lds, int C, const int32_t *__restrict__ idx, int64_t idx_ld,
                  int64_t K, int k, const float *__restrict__ p, float eps, float *__restrict__ L, int64_t ldl,
                  int n_groups) {
    pdl_enter();
    static_assert(U == 8, "two groups of four ranks per iteration");
    constexpr int NPB = kAccumThreads / TPN;
    // the gathered rows of S are re-read ~k*K/N (33 at c4) times: ask L2 to keep them (evict-last), measured
    // 1.15 -> 0.80 ms at c4
    const uint64_t keep = l2_policy_evict_last();
    // dynamic shared memory, sized to k (a few KB at k = 100, so a CTA also fits beside other resident kernels):
    // BYTE offsets idx * lds * 4 per neuron (host: N * lds < 2^30), then the rank weights p and the FMA constants c
    extern __shared__ __align__(16) unsigned char accum_smem[];
    const int kpad = (k + 3) & ~3;
    uint32_t *s_off = reinterpret_cast<uint32_t *>(accum_smem);          // [NPB][kpad]
    float *s_p = reinterpret_cast<float *>(s_off + NPB * kpad);           // [kpad]
    float *s_c = s_p + kpad;                                              // [kpad]

    const int tile = blockIdx.x / n_groups;
    const int group = blockIdx.x - tile * n_groups;
    const int64_t j0 = int64_t(group) * NPB;
    const int tid = threadIdx.x;

    for (int n = 0; n < NPB; ++n) {
        const int64_t j = j0 + n;
        for (int r = tid; r < k; r += kAccumThreads)
            s_off[n * kpad + r] = j < K ? uint32_t(idx[int64_t(r) * idx_ld + j]) * uint32_t(lds) * 4u : 0u;
    }
    bool p_ok = true;
    if (SOFT)
        for (int r = tid; r < k; r += kAccumThreads) {
            const float w = p[r];
            s_p[r] = w;
            if (FUSED) s_c[r] = static_cast<float>(1.0 - double(w) + double(eps));
            p_ok = p_ok && (w >= 0.f) && (w <= 1.f);
        }
    const bool grouped = GROUPED && (__syncthreads_and(p_ok) != 0);
    if (!GROUPED) __syncthreads();

    const int n = tid / TPN;
    const int tl = tid - n * TPN;
    const int64_t j = j0 + n;
    const int c0 = (tile * TPN + tl) * 4;
    if (j >= K || c0 >= C) return;
    const int nvalid = min(4, C - c0);
    // a vector load may run past C only inside the row's padding (c0 + 4 <= lds is guaranteed by
    // the host when VEC); the scalar path loads exactly the valid columns
    const char *base = reinterpret_cast<const char *>(S + c0);
    const uint32_t *my_off = s_off + n * kpad;

    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    int r = 0;
    if (grouped) {
        for (; r + U <= k; r += U) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = load_row<VEC>(base, my_off[r + u], nvalid, keep);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float *acc = h ? acc1 : acc0;
                float w[4], c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    w[u] = SOFT ? s_p[r + 4 * h + u] : 0.f;
                    c[u] = SOFT && FUSED ? s_c[r + 4 * h + u] : 0.f;
                }
                const float4 *q = s + 4 * h;
                acc[0] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].x, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].x, w[1], c[1], eps)) *
                                        (term<SOFT, FUSED>(q[2].x, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].x, w[3], c[3], eps)));
                acc[1] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].y, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].y, w[1], c[1], eps)) *
                                        (term<SOFT, FUSED>(q[2].y, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].y, w[3], c[3], eps)));
                acc[2] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].z, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].z, w[1], c[1], eps)) *
                                        (term<SOFT, FUSED>(q[2].z, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].z, w[3], c[3], eps)));
                acc[3] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].w, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].w, w[1], c[1], eps)) *
                                        (term<SOFT, FUSED>(q[2].w, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].w, w[3], c[3], eps)));
            }
        }
        for (; r + 4 <= k; r += 4) {
            float4 q[4];
            float w[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q[u] = load_row<VEC>(base, my_off[r + u], nvalid, keep);
                w[u] = SOFT ? s_p[r + u] : 0.f;
                c[u] = SOFT && FUSED ? s_c[r + u] : 0.f;
            }
            acc0[0] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].x, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].x, w[1], c[1], eps)) *
                                     (term<SOFT, FUSED>(q[2].x, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].x, w[3], c[3], eps)));
            acc0[1] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].y, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].y, w[1], c[1], eps)) *
                                     (term<SOFT, FUSED>(q[2].y, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].y, w[3], c[3], eps)));
            acc0[2] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].z, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].z, w[1], c[1], eps)) *
                                     (term<SOFT, FUSED>(q[2].z, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].z, w[3], c[3], eps)));
            acc0[3] += lg2_fast<FTZ>((term<SOFT, FUSED>(q[0].w, w[0], c[0], eps) * term<SOFT, FUSED>(q[1].w, w[1], c[1], eps)) *
                                     (term<SOFT, FUSED>(q[2].w, w[2], c[2], eps) * term<SOFT, FUSED>(q[3].w, w[3], c[3], eps)));
        }
    } else {
        for (; r + U <= k; r += U) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = load_row<VEC>(base, my_off[r + u], nvalid, keep);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float w = SOFT ? s_p[r + u] : 0.f;
                float *acc = (u & 1) ? acc1 : acc0;
                acc[0] += lg2_fast<FTZ>(term<SOFT, false>(s[u].x, w, 0.f, eps));
                acc[1] += lg2_fast<FTZ>(term<SOFT, false>(s[u].y, w, 0.f, eps));
                acc[2] += lg2_fast<FTZ>(term<SOFT, false>(s[u].z, w, 0.f, eps));
                acc[3] += lg2_fast<FTZ>(term<SOFT, false>(s[u].w, w, 0.f, eps));
            }
        }
    }
    for (; r < k; ++r) {
        const float4 s = load_row<VEC>(base, my_off[r], nvalid, keep);
        const float w = SOFT ? s_p[r] : 0.f;
        if (SOFT && FUSED && grouped) {
            const float c = s_c[r];
            acc0[0] += lg2_fast<FTZ>(term<SOFT, true>(s.x, w, c, eps));
            acc0[1] += lg2_fast<FTZ>(term<SOFT, true>(s.y, w, c, eps));
            acc0[2] += lg2_fast<FTZ>(term<SOFT, true>(s.z, w, c, eps));
            acc0[3] += lg2_fast<FTZ>(term<SOFT, true>(s.w, w, c, eps));
        } else {
            acc0[0] += lg2_fast<FTZ>(term<SOFT, false>(s.x, w, 0.f, eps));
            acc0[1] += lg2_fast<FTZ>(term<SOFT, false>(s.y, w, 0.f, eps));
            acc0[2] += lg2_fast<FTZ>(term<SOFT, false>(s.z, w, 0.f, eps));
            acc0[3] += lg2_fast<FTZ>(term<SOFT, false>(s.w, w, 0.f, eps));
        }
    }
    constexpr float kLn2 = 0.693147180559945309417f;
    float *out = L + j * ldl + c0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (e < nvalid) out[e] = (acc0[e] + acc1[e]) * kLn2;
}

template <int TPN, bool SOFT, bool VEC>
static int launch_accum(const float *S, int64_t lds, int C, const int32_t *idx, int64_t idx_ld, int64_t K, int k,
                        const float *p, float eps, float *L, int64_t ldl, cudaStream_t st, bool probabilities, size_t pad_smem) {
    const bool ftz = eps >= 1.17549435e-38f;
    constexpr int NPB = kAccumThreads / TPN;
    const int n_tiles = ceil_div(C, TPN * 4);
    const int64_t n_groups = ceil_div<int64_t>(K, NPB);
    const int64_t blocks = n_groups * n_tiles;
    if (blocks > 0x7FFFFFFFll) return MCD_ERR_UNSUPPORTED;
    // grouped logs need eps >= 1e-9 (see the kernel).  Tunable accum_unroll: 0 = grouped, term as one FMA (default);
    // 2 = grouped, term in the reference's operation order; 1 = one lg2 per term, reference order
    const int mode = static_cast<int>(tunable(kAccumUnroll));
    // ... and only for a caller that vouches for S in [0, 1] (mcd_wpmi_accum_prob_f32, the fused calls): with a negative
    // entry two negative terms would multiply to a positive product and the log would come out finite instead of NaN
    const bool grouped = probabilities && eps >= 1e-9f && eps <= 1.0f && mode != 1;
    const unsigned nb = static_cast<unsigned>(blocks);
    // pad_smem: unused shared memory that caps how many of these CTAs fit beside another resident kernel (abi.cu)
    const size_t sm = size_t(NPB + 2) * size_t((k + 3) & ~3) * 4 + pad_smem;
    const int ng = static_cast<int>(n_groups);
    if (grouped && mode != 2)
        launch_pdl((wpmi_accum_kernel<TPN, 8, SOFT, VEC, true, true, true>), dim3(nb), dim3(kAccumThreads), sm, st, S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, ng);
    else if (grouped)
        launch_pdl((wpmi_accum_kernel<TPN, 8, SOFT, VEC, true, true, false>), dim3(nb), dim3(kAccumThreads), sm, st, S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, ng);
    else if (ftz)
        launch_pdl((wpmi_accum_kernel<TPN, 8, SOFT, VEC, true, false, false>), dim3(nb), dim3(kAccumThreads), sm, st, S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, ng);
    else
        launch_pdl((wpmi_accum_kernel<TPN, 8, SOFT, VEC, false, false, false>), dim3(nb), dim3(kAccumThreads), sm, st, S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, ng);
    return check_launch();
}

template <bool SOFT, bool VEC>
static int dispatch_tile(int tpn, const float *S, int64_t lds, int C, const int32_t *idx, int64_t idx_ld, int64_t K, int k,
                         const float *p, float eps, float *L, int64_t ldl, cudaStream_t st, bool probabilities,
                         size_t pad_smem) {
    switch (tpn) {
        case 192: return launch_accum<192, SOFT, VEC>(S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, st, probabilities, pad_smem);
        case 96: return launch_accum<96, SOFT, VEC>(S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, st, probabilities, pad_smem);
        case 64: return launch_accum<64, SOFT, VEC>(S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, st, probabilities, pad_smem);
        case 48: return launch_accum<48, SOFT, VEC>(S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, st, probabilities, pad_smem);
        case 16: return launch_accum<16, SOFT, VEC>(S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, st, probabilities, pad_smem);
        default: return launch_accum<32, SOFT, VEC>(S, lds, C, idx, idx_ld, K, k, p, eps, L, ldl, st, probabilities, pad_smem);
    }
}

int wpmi_accum_range(const float *S, int64_t lds, int64_t N, int64_t C, const int32_t *idx, int64_t idx_ld, int64_t K,
                     int64_t k, const float *p, float min_prob, float *L, int64_t ldl, cudaStream_t st, bool probabilities,
                     size_t pad_smem) {
    if (!S || !idx || !L || N < 1 || C < 1 || K < 1 || k < 1 || lds < C || ldl < C || idx_ld < K) return MCD_ERR_INVALID_ARGUMENT;
    if (k > kAccumMaxK || C > (1 << 24) || N * lds >= (int64_t(1) << 30)) return MCD_ERR_UNSUPPORTED;
    // 16-byte loads need aligned rows and must stay inside a row (padding included)
    const bool vec = (lds % 4 == 0) && (reinterpret_cast<uintptr_t>(S) % 16 == 0) && (ceil_div<int64_t>(C, 4) * 4 <= lds);
    // threads per neuron: tunable "accum_tile" = concepts per tile (multiple of 4).  Default: 128-concept tiles
    // (512-byte row pieces) once S no longer fits in L2 -- measured 12 % faster than whole rows at N = 100k,
    // C = 763 because a tile's slice of S (51 MB) stays L2-resident while all neurons gather from it.
    int64_t tile_c = tunable(kAccumTile);
    int tpn;
    if (tile_c <= 0) tile_c = (N * lds * 4 > (int64_t(96) << 20) && C > 128) ? 128 : C;
    const int64_t want = ceil_div<int64_t>(tile_c, 4);
    if (want > 96) tpn = 192;
    else if (want > 64) tpn = 96;
    else if (want > 48) tpn = 64;
    else if (want > 32) tpn = 48;
    else if (want > 16) tpn = 32;
    else tpn = 16;
    const int Ci = static_cast<int>(C), ki = static_cast<int>(k);
    if (p) return vec ? dispatch_tile<true, true>(tpn, S, lds, Ci, idx, idx_ld, K, ki, p, min_prob, L, ldl, st, probabilities, pad_smem)
                      : dispatch_tile<true, false>(tpn, S, lds, Ci, idx, idx_ld, K, ki, p, min_prob, L, ldl, st, probabilities, pad_smem);
    return vec ? dispatch_tile<false, true>(tpn, S, lds, Ci, idx, idx_ld, K, ki, p, min_prob, L, ldl, st, probabilities, pad_smem)
               : dispatch_tile<false, false>(tpn, S, lds, Ci, idx, idx_ld, K, ki, p, min_prob, L, ldl, st, probabilities, pad_smem);
}

}  // namespace mcd

extern "C" int mcd_wpmi_accum_f32(const float *S, int64_t lds, int64_t N, int64_t C, const int32_t *idx, int64_t K,
                                  int64_t k, const float *p, float min_prob, float *L, int64_t ldl,
                                  mcd_stream_t stream) {
    return mcd::wpmi_accum_range(S, lds, N, C, idx, K, K, k, p, min_prob, L, ldl, static_cast<cudaStream_t>(stream), false, 0);
}

extern "C" int mcd_wpmi_accum_prob_f32(const float *S, int64_t lds, int64_t N, int64_t C, const int32_t *idx, int64_t K,
                                       int64_t k, const float *p, float min_prob, float *L, int64_t ldl,
                                       mcd_stream_t stream) {
    return mcd::wpmi_accum_range(S, lds, N, C, idx, K, K, k, p, min_prob, L, ldl, static_cast<cudaStream_t>(stream), true, 0);
}
