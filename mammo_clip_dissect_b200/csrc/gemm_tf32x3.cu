// K1 on the 5th-generation tensor cores: P = normalise(I) * normalise(T)^T   (reference concept_vit/utils.py:577-594)
//
// The reference multiplies in true fp32 (torch's allow_tf32 is off) and soft_wpmi then scales the logits by a = 10,
// so a single-pass TF32 product (6.5e-4 relative error on the softmax) is not acceptable.  The product
// is therefore evaluated as a 3-term split on tcgen05 kind::tf32 with fp32 accumulation in TMEM:
//        x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exactly representable in TF32),  lo = x - hi
//        I T^T  ~=  Ihi Thi^T + Ihi Tlo^T + Ilo Thi^T                                  (the lo*lo term is < 2^-20)
// which reproduces fp32-grade results (1.4e-6 on the softmax in the survey's emulation; measured in the tests).
//
//   prepare_rows_kernel   one warp per row: ||x||, x/||x|| with the reference's per-element division, hi / lo split,
//                         written as two row-padded fp32 matrices (rows to a multiple of the tile, D to 32)
//   gemm_tf32x3_kernel    CTA = one 128 x 128 output tile.  warp 0: TMA producer (4 operand tiles per 32-wide
//                         k-block, 128-byte swizzle, 3-stage mbarrier ring); warp 1: one thread issues 12
//                         tcgen05.mma (M128 N128 K8) per k-block and commits to the stage's `empty` barrier;
//                         warps 2-5: epilogue, tcgen05.ld 32 lanes x 32 columns at a time -> global stores.
//                         The N-tile index is the fastest grid dimension so the CTAs sharing an I tile hit L2.
// The tensor core's fp32 accumulation truncates, so its error grows linearly with the length of the k loop
// (measured 5.4e-6 of max|P| at D = 512 with one accumulator vs 1e-6 for the fp32 FFMA kernel).  The k-blocks are
// therefore dealt round-robin onto 4 independent TMEM accumulators (4 x 128 columns = the whole TMEM) that the
// epilogue adds with round-to-nearest fp32 adds.
#include <cuda.h>
#include <cstring>

#include "common.cuh"

namespace mcd {

constexpr int kBM = 128, kBN = 128, kBK = 32;            // CTA tile; kBK fp32 = one 128-byte swizzle row
constexpr int kGemmStages = 3;
constexpr int kAccSegs = 4;                              // independent TMEM accumulators (k-blocks round-robin)
constexpr int kTmemCols = kAccSegs * kBN;                // 512 = all of TMEM
constexpr int kTcThreads = 192;                          // producer warp, MMA warp, 4 epilogue warps
constexpr uint32_t kATileBytes = kBM * kBK * 4;          // 16 KB
constexpr uint32_t kBTileBytes = kBN * kBK * 4;          // 16 KB
constexpr uint32_t kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;   // hi + lo of both operands: 64 KB
constexpr size_t kTcSmemBytes = size_t(kGemmStages) * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(256)
prepare_rows_kernel(const float *__restrict__ X, int64_t ldx, int64_t R, int64_t D, int normalize,
                    float *__restrict__ Xhi, float *__restrict__ Xlo, int64_t Rpad, int64_t Dpad) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (r >= Rpad) return;
    float *hi = Xhi + r * Dpad, *lo = Xlo + r * Dpad;
    if (r >= R) {
        for (int64_t d = lane; d < Dpad; d += 32) hi[d] = lo[d] = 0.f;
        return;
    }
    const float *x = X + r * ldx;
    float nrm = 1.f;
    if (normalize) {
        float s = 0.f;
        for (int64_t d = lane; d < D; d += 32) {
            const float v = x[d];
            s = fmaf(v, v, s);
        }
        nrm = sqrtf(warp_sum(s));
    }
    for (int64_t d = lane; d < Dpad; d += 32) {
        float v = 0.f;
        if (d < D) v = normalize ? __fdiv_rn(x[d], nrm) : x[d];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[d] = h;
        lo[d] = __fsub_rn(v, h);          // exact
    }
}

// ---- tcgen05 / TMA plumbing -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    // a mis-programmed pipeline must fail loudly, never hang the GPU
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO); version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kBN >> 3) << 17) | (uint32_t(kBM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kIdescTf32), "r"(uint32_t(accumulate))
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the load alone (asynchronous until tmem_ld_wait): several can be in flight
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                   int64_t M, int64_t Nn, int num_kb, float *__restrict__ P, int64_t ldp) {
    pdl_enter();
    extern __shared__ unsigned char smem_dyn[];
    // 128-byte swizzle needs 1024-byte aligned tiles
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *aligned = smem_dyn + (base - smem_u32(smem_dyn));
    uint64_t *full = reinterpret_cast<uint64_t *>(aligned + size_t(kGemmStages) * kStageBytes);
    uint64_t *empty = full + kGemmStages;
    uint64_t *tmem_full = empty + kGemmStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {     // TMEM: 4 x 128 fp32 accumulator columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kGemmStages, use = kb / kGemmStages;
                if (use > 0) mbar_wait_bounded(&empty[s], (use - 1) & 1);
                mbar_arrive_expect_tx(&full[s], kStageBytes);
                const uint32_t st = base + s * kStageBytes;
                const int kx = kb * kBK;
                tma_load_2d(st, &mapAhi, kx, m_tile * kBM, &full[s]);
                tma_load_2d(st + kATileBytes, &mapAlo, kx, m_tile * kBM, &full[s]);
                tma_load_2d(st + 2 * kATileBytes, &mapBhi, kx, n_tile * kBN, &full[s]);
                tma_load_2d(st + 2 * kATileBytes + kBTileBytes, &mapBlo, kx, n_tile * kBN, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kGemmStages, use = kb / kGemmStages;
                mbar_wait_bounded(&full[s], use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * kStageBytes;
                const uint64_t ahi = umma_desc_k128(st), alo = umma_desc_k128(st + kATileBytes);
                const uint64_t bhi = umma_desc_k128(st + 2 * kATileBytes), blo = umma_desc_k128(st + 2 * kATileBytes + kBTileBytes);
                const uint32_t acc = tmem_base + uint32_t(kb % kAccSegs) * kBN;     // this k-block's accumulator
                for (int kk = 0; kk < kBK / 8; ++kk) {          // K = 8 per tf32 MMA: 32 bytes inside the swizzle row
                    const uint64_t adv = uint64_t((kk * 32) >> 4);
                    umma_tf32(acc, ahi + adv, bhi + adv, kb >= kAccSegs || kk > 0);
                    umma_tf32(acc, ahi + adv, blo + adv, true);
                    umma_tf32(acc, alo + adv, bhi + adv, true);
                }
                umma_commit(&empty[s]);                          // smem stage reusable once these MMAs have read it
            }
            umma_commit(tmem_full);                              // accumulator complete
        }
    } else {
        // epilogue warps 2..5: warp w may touch TMEM lanes 32*(w%4) .. +31 only
        const int quarter = warp & 3;
        mbar_wait_bounded(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t row = int64_t(m_tile) * kBM + quarter * 32 + lane;
        const int64_t col0 = int64_t(n_tile) * kBN;
        const bool vec_ok = (ldp % 4 == 0) && (reinterpret_cast<uintptr_t>(P) % 16 == 0);
        const int nseg = num_kb < kAccSegs ? num_kb : kAccSegs;
#pragma unroll 1
        for (int c = 0; c < kBN; c += 32) {
            float v[32];
            const uint32_t t0 = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c);
            if (nseg == kAccSegs) {          // all four accumulators of this column chunk in flight, one wait
                uint32_t r[kAccSegs][32];
#pragma unroll
                for (int sgm = 0; sgm < kAccSegs; ++sgm) tmem_ld_32x32_issue(t0 + uint32_t(sgm * kBN), r[sgm]);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = __uint_as_float(r[0][i]);
#pragma unroll
                    for (int sgm = 1; sgm < kAccSegs; ++sgm) v[i] = __fadd_rn(v[i], __uint_as_float(r[sgm][i]));
                }
            } else {
                tmem_ld_32x32(t0, v);
                for (int sgm = 1; sgm < nseg; ++sgm) {
                    float w[32];
                    tmem_ld_32x32(t0 + uint32_t(sgm * kBN), w);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __fadd_rn(v[i], w[i]);
                }
            }
            if (row < M) {
                float *dst = P + row * ldp + col0 + c;
                if (vec_ok && col0 + c + 32 <= Nn) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + c + i < Nn) dst[i] = v[i];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---- the same GEMM with the row softmax fused in: one CTA per 128-row band -------------------------------------------
// A CTA walks all N-tiles of its band (6 for the 763-concept set), so the producer keeps prefetching operands of the
// next tile while the epilogue drains the accumulators of the current one, and an epilogue thread -- which owns one
// row of the band, TMEM lane = row -- sees every logit of its row: it keeps an online (max, sum of exponentials) pair
// while it stores P.  When the last tile is done the four epilogue warps rescale the band row by row with coalesced
// accesses (the band's P, 390 KB, is still in L2):  S = exp(a*P - max) / sum, per element the reference's operation
// sequence (separately rounded a*P, expf, true division).  Only the order of the row sum differs from the stand-alone
// kernel (softmax_rows.cu), by a few ulp.
constexpr int kBandEpiThreads = 128;

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_band_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                        const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                        int64_t M, int64_t Nn, int num_kb, int n_tiles, float *__restrict__ P, int64_t ldp,
                        float *__restrict__ S, int64_t lds, float a) {
    pdl_enter();
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *aligned = smem_dyn + (base - smem_u32(smem_dyn));
    uint64_t *full = reinterpret_cast<uint64_t *>(aligned + size_t(kGemmStages) * kStageBytes);
    uint64_t *empty = full + kGemmStages;
    uint64_t *tmem_full = empty + kGemmStages;
    uint64_t *tmem_empty = tmem_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);
    __shared__ float row_max[kBM], row_sum[kBM];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);            // one arrival per epilogue warp
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;                       // k-block counter over all tiles: ring position and phase
            for (int nt = 0; nt < n_tiles; ++nt)
                for (int kb = 0; kb < num_kb; ++kb, ++g) {
                    const int s = g % kGemmStages, use = g / kGemmStages;
                    if (use > 0) mbar_wait_bounded(&empty[s], (use - 1) & 1);
                    mbar_arrive_expect_tx(&full[s], kStageBytes);
                    const uint32_t st = base + s * kStageBytes;
                    const int kx = kb * kBK;
                    tma_load_2d(st, &mapAhi, kx, m_tile * kBM, &full[s]);
                    tma_load_2d(st + kATileBytes, &mapAlo, kx, m_tile * kBM, &full[s]);
                    tma_load_2d(st + 2 * kATileBytes, &mapBhi, kx, nt * kBN, &full[s]);
                    tma_load_2d(st + 2 * kATileBytes + kBTileBytes, &mapBlo, kx, nt * kBN, &full[s]);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int g = 0;
            for (int nt = 0; nt < n_tiles; ++nt) {
                if (nt > 0) {                // the epilogue has read the previous tile's accumulators
                    mbar_wait_bounded(tmem_empty, (nt - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kb = 0; kb < num_kb; ++kb, ++g) {
                    const int s = g % kGemmStages, use = g / kGemmStages;
                    mbar_wait_bounded(&full[s], use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * kStageBytes;
                    const uint64_t ahi = umma_desc_k128(st), alo = umma_desc_k128(st + kATileBytes);
                    const uint64_t bhi = umma_desc_k128(st + 2 * kATileBytes), blo = umma_desc_k128(st + 2 * kATileBytes + kBTileBytes);
                    const uint32_t acc = tmem_base + uint32_t(kb % kAccSegs) * kBN;
                    for (int kk = 0; kk < kBK / 8; ++kk) {
                        const uint64_t adv = uint64_t((kk * 32) >> 4);
                        umma_tf32(acc, ahi + adv, bhi + adv, kb >= kAccSegs || kk > 0);
                        umma_tf32(acc, ahi + adv, blo + adv, true);
                        umma_tf32(acc, alo + adv, bhi + adv, true);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(tmem_full);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int r_in_band = quarter * 32 + lane;
        const int64_t row = int64_t(m_tile) * kBM + r_in_band;
        const bool vec_ok = (ldp % 4 == 0) && (reinterpret_cast<uintptr_t>(P) % 16 == 0);
        const int nseg = num_kb < kAccSegs ? num_kb : kAccSegs;
        float m = -INFINITY, ssum = 0.f;
#pragma unroll 1
        for (int nt = 0; nt < n_tiles; ++nt) {
            mbar_wait_bounded(tmem_full, nt & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t col0 = int64_t(nt) * kBN;
#pragma unroll 1
            for (int c = 0; c < kBN; c += 32) {
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c), v);
                for (int sgm = 1; sgm < nseg; ++sgm) {
                    float w[32];
                    tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(sgm * kBN + c), w);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __fadd_rn(v[i], w[i]);
                }
                if (row < M && col0 + c < Nn) {
                    float *dst = P + row * ldp + col0 + c;
                    const bool whole = col0 + c + 32 <= Nn;
                    if (vec_ok && whole) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + c + i < Nn) dst[i] = v[i];
                    }
                    if (S != nullptr) {
                        float z[32], mx = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            z[i] = (whole || col0 + c + i < Nn) ? __fmul_rn(a, v[i]) : -INFINITY;
                            mx = fmaxf(mx, z[i]);
                        }
                        if (mx > m) {
                            ssum *= (m == -INFINITY) ? 0.f : expf(m - mx);
                            m = mx;
                        }
                        const float ms = (m == -INFINITY) ? 0.f : m;
#pragma unroll
                        for (int i = 0; i < 32; ++i) ssum += expf(__fsub_rn(z[i], ms));     // exp(-inf) = 0 for the padding
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
        }
        if (S != nullptr) {
            row_max[r_in_band] = m;
            row_sum[r_in_band] = ssum;
            // the P values below were written by other threads of these four warps: make them visible, then meet
            __threadfence_block();
            asm volatile("bar.sync 1, %0;" ::"n"(kBandEpiThreads) : "memory");
            for (int r = quarter; r < kBM; r += 4) {        // warp per row, coalesced
                const int64_t grow = int64_t(m_tile) * kBM + r;
                if (grow >= M) break;
                const float rm = row_max[r], rs = row_sum[r];
                const float *src = P + grow * ldp;
                float *dst = S + grow * lds;
                for (int64_t c = lane; c < lds; c += 32)
                    dst[c] = c < Nn ? __fdiv_rn(expf(__fsub_rn(__fmul_rn(a, src[c]), rm)), rs) : 0.f;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---- cos / cos^3 similarities on the tensor cores (reference concept_vit/similarity.py:7-47) ---------------------------
// out[j, c] = sum_i f(A[i, j]) * f(P[i, c]): the contraction runs over the N probe images, the slow axis of both
// row-major inputs.  prepare_cols_kernel applies f (centre, cube, divide by the clipped norm -- the reference's fp32
// operation sequence -- or divide by the norm), splits the result into hi + lo and writes it TRANSPOSED ([columns, images],
// zero-padded to the tile grid), which makes both operands K-major for the same UMMA descriptors as K1.
// gemm_tf32x3_long_kernel is K1's kernel for a long contraction: the tensor core's fp32 accumulation truncates, so an
// accumulator is only trusted for kFlushBlocks k-blocks (48 MMAs, the same depth as K1 at D = 512); the k-blocks of a
// group alternate between the two accumulators of a TMEM set, and while the MMA warp fills the other set the four
// epilogue warps add the finished set into fp32 registers with round-to-nearest adds (a thread owns one output row:
// 128 running sums).  The error therefore no longer grows with N.
enum { kPrepCos = 1, kPrepCos3 = 2 };

// Column statistics in two row-parallel passes (the thread-per-column form of dense_sim.cu leaves all but a few SMs idle
// for the 763-concept matrix): grid (column groups of 32, row chunks); a block reduces its chunk over 8 row lanes and
// writes one partial per column, part[chunk][column]; consumers add the chunks in order (fixed order: deterministic).
//   STAT 0: sum x                      STAT 1: sum x^2                      STAT 2: sum ((x - mean)^3)^2
struct ColSrc {
    const float *X;            // [N, M] row-major
    int64_t ldx, M;
    const float *sum_part;     // STAT 2: the folded column sums (row 0)
    float *part;               // [chunks][M]
};

// both matrices of the call in one launch: blockIdx.z picks the source (the narrower one leaves its surplus CTAs idle)
template <int STAT>
__global__ void __launch_bounds__(256)
col_partial_kernel(ColSrc s0, ColSrc s1, int64_t N, int64_t rows_per_chunk) {
    pdl_enter();
    __shared__ float red[8][33];
    const ColSrc &src = blockIdx.z == 0 ? s0 : s1;
    const float *__restrict__ X = src.X;
    const int64_t ldx = src.ldx, M = src.M;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t m = int64_t(blockIdx.x) * 32 + cx;
    if (int64_t(blockIdx.x) * 32 >= M) return;
    const int64_t r0 = int64_t(blockIdx.y) * rows_per_chunk, r1 = min(N, r0 + rows_per_chunk);
    float mean = 0.f;
    if (STAT == 2 && m < M) mean = (0.f + src.sum_part[m]) / static_cast<float>(N);
    float s = 0.f;
    if (m < M)
        for (int64_t i = r0 + ry; i < r1; i += 8) {
            float v = X[i * ldx + m];
            if (STAT == 2) {
                const float d = __fsub_rn(v, mean);
                v = __fmul_rn(__fmul_rn(d, d), d);
            }
            s = STAT == 0 ? s + v : fmaf(v, v, s);
        }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && m < M) {
        float t = red[0][cx];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += red[q][cx];
        src.part[int64_t(blockIdx.y) * M + m] = t;
    }
}

// part[0][m] = sum over the chunks of part[q][m], in chunk order (the sum every consumer used to redo per CTA: the 157
// row-tile CTAs of a column block each re-added 37 partials per column before touching their 1024 elements -- 34 us for a
// 31 MB transpose at c5).  Consumers then read one row: 0 + total == total, the same bits as before.  Both matrices in one
// launch (blockIdx.y).
__global__ void __launch_bounds__(256)
col_fold_kernel(float *__restrict__ part0, int64_t M0, float *__restrict__ part1, int64_t M1, int n_chunks) {
    pdl_enter();
    float *part = blockIdx.y == 0 ? part0 : part1;
    const int64_t M = blockIdx.y == 0 ? M0 : M1;
    const int64_t m = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (m >= M) return;
    float t = 0.f;
    for (int q = 0; q < n_chunks; ++q) t += part[int64_t(q) * M + m];
    part[m] = t;
}

template <int MODE>
__global__ void __launch_bounds__(256)
prepare_cols_kernel(const float *__restrict__ X, int64_t ldx, int64_t N, int64_t M, const float *__restrict__ sum_part,
                    const float *__restrict__ sq_part, int n_chunks, int64_t part_ld, float min_norm,
                    float *__restrict__ Thi, float *__restrict__ Tlo, int64_t Npad) {
    pdl_enter();
    __shared__ float s_hi[32][33], s_lo[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t m0 = int64_t(blockIdx.x) * 32, i0 = int64_t(blockIdx.y) * 32;
    const int64_t m = m0 + tx;
    float mu = 0.f, nr = 1.f;
    if (m < M) {
        float t = 0.f, q2 = 0.f;
        for (int q = 0; q < n_chunks; ++q) {
            if (MODE == kPrepCos3) t += sum_part[int64_t(q) * part_ld + m];
            q2 += sq_part[int64_t(q) * part_ld + m];
        }
        mu = t / static_cast<float>(N);
        nr = sqrtf(q2);
        if (MODE == kPrepCos3) {
            nr = fmaxf(nr, min_norm);            // torch.clip(norm, min_norm); NaN stays NaN
            if (q2 != q2) nr = q2;
        }
    }
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = i0 + r;
        float v = 0.f;
        if (m < M && i < N) {
            const float x = X[i * ldx + m];
            if (MODE == kPrepCos3) {
                const float d = __fsub_rn(x, mu);
                v = __fdiv_rn(__fmul_rn(__fmul_rn(d, d), d), nr);
            } else {
                v = __fdiv_rn(x, nr);
            }
        }
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        s_hi[r][tx] = h;
        s_lo[r][tx] = __fsub_rn(v, h);           // exact (NaN / inf stay NaN / inf in hi; lo becomes NaN: propagates)
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t o = (m0 + r) * Npad + i0 + tx;
        Thi[o] = s_hi[tx][r];
        Tlo[o] = s_lo[tx][r];
    }
}

// out = sum of the split-K partial tiles, in split order
__global__ void __launch_bounds__(256)
splitk_combine_kernel(const float *__restrict__ part, int splits, int64_t M, int64_t Nn, int64_t ldpart, int64_t plane,
                      float *__restrict__ out, int64_t ldo) {
    pdl_enter();
    const int64_t n = int64_t(blockIdx.x) * 256 + threadIdx.x, m = blockIdx.y;
    if (n >= Nn) return;
    float t = part[m * ldpart + n];
    for (int s = 1; s < splits; ++s) t = __fadd_rn(t, part[int64_t(s) * plane + m * ldpart + n]);
    out[m * ldo + n] = t;
}

constexpr int kFlushBlocks = 8;                          // k-blocks per accumulator group (2 accumulators x 4 k-blocks)

// One CTA = `tiles_per_cta` consecutive 128-column output tiles of one 128-row band, processed back to back through the
// same pipelines: the producer's ring and the two TMEM sets never drain between tiles, so the epilogue of a tile (TMEM ->
// registers -> global, 64 KB of stores) runs under the MMAs of the next one.  Within a tile the k-blocks go in groups
// of kFlushBlocks onto alternating TMEM sets; the epilogue adds every finished group into 128 fp32 registers per thread
// (round-to-nearest adds), which bounds the tensor core's truncating accumulation to 4 k-blocks per accumulator
// whatever the contraction length.
//   cos / cos^3 : tiles_per_cta = 1, the contraction (probe images) split over blockIdx.z, outputs are split-K planes;
//   K1          : tiles_per_cta = all column tiles of the band, no split.  An epilogue thread owns one output row, so it
//                 can keep the row's online softmax pair (max, sum of exponentials of a * value) while the tiles go
//                 by: row_stats [M][2] feeds softmax_from_stats_kernel.
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_long_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                        const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                        int64_t M, int64_t Nn, int total_kb, int kb_per_split, float *__restrict__ Cout, int64_t ldc,
                        int64_t split_plane, int tiles_per_cta, int n_tiles_total, float a, float *__restrict__ row_stats,
                        int terms) {
    pdl_enter();
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *aligned = smem_dyn + (base - smem_u32(smem_dyn));
    uint64_t *full = reinterpret_cast<uint64_t *>(aligned + size_t(kGemmStages) * kStageBytes);
    uint64_t *empty = full + kGemmStages;
    uint64_t *tmem_full = empty + kGemmStages;           // [2]: one per TMEM set
    uint64_t *tmem_empty = tmem_full + 2;                // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile0 = blockIdx.x * tiles_per_cta, m_tile = blockIdx.y;
    const int ntiles = min(tiles_per_cta, n_tiles_total - n_tile0);
    // split-K: blockIdx.z takes k-blocks [kb_first, kb_first + num_kb) and writes its own output plane
    const int kb_first = blockIdx.z * kb_per_split;
    const int num_kb = min(kb_per_split, total_kb - kb_first);
    Cout += int64_t(blockIdx.z) * split_plane;
    const int ngroups = (num_kb + kFlushBlocks - 1) / kFlushBlocks;          // per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);        // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {     // TMEM: 2 sets x 2 accumulators x 128 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int f = 0;                           // k-block counter over all tiles: ring position and phase
            for (int t = 0; t < ntiles; ++t)
                for (int kb = 0; kb < num_kb; ++kb, ++f) {
                    const int s = f % kGemmStages, use = f / kGemmStages;
                    if (use > 0) mbar_wait_bounded(&empty[s], (use - 1) & 1);
                    mbar_arrive_expect_tx(&full[s], kStageBytes);
                    const uint32_t st = base + s * kStageBytes;
                    const int kx = (kb_first + kb) * kBK;
                    tma_load_2d(st, &mapAhi, kx, m_tile * kBM, &full[s]);
                    tma_load_2d(st + kATileBytes, &mapAlo, kx, m_tile * kBM, &full[s]);
                    tma_load_2d(st + 2 * kATileBytes, &mapBhi, kx, (n_tile0 + t) * kBN, &full[s]);
                    tma_load_2d(st + 2 * kATileBytes + kBTileBytes, &mapBlo, kx, (n_tile0 + t) * kBN, &full[s]);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int f = 0;
            for (int t = 0; t < ntiles; ++t)
                for (int kb = 0; kb < num_kb; ++kb, ++f) {
                    const int g = kb / kFlushBlocks, in_g = kb - g * kFlushBlocks;
                    const int G = t * ngroups + g, set = G & 1;              // group counter over all tiles
                    if (in_g == 0 && G >= 2) {               // the epilogue has drained this set's previous group
                        mbar_wait_bounded(&tmem_empty[set], ((G >> 1) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const int s = f % kGemmStages, use = f / kGemmStages;
                    mbar_wait_bounded(&full[s], use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * kStageBytes;
                    const uint64_t ahi = umma_desc_k128(st), alo = umma_desc_k128(st + kATileBytes);
                    const uint64_t bhi = umma_desc_k128(st + 2 * kATileBytes), blo = umma_desc_k128(st + 2 * kATileBytes + kBTileBytes);
                    const uint32_t acc = tmem_base + uint32_t(set * 2 + (in_g & 1)) * kBN;
                    for (int kk = 0; kk < kBK / 8; ++kk) {
                        const uint64_t adv = uint64_t((kk * 32) >> 4);
                        umma_tf32(acc, ahi + adv, bhi + adv, in_g >= 2 || kk > 0);      // first use of the accumulator in its group: overwrite
                        if (terms >= 3) {            // (terms = 1: measurement aid only, plain TF32)
                            umma_tf32(acc, ahi + adv, blo + adv, true);
                            umma_tf32(acc, alo + adv, bhi + adv, true);
                        }
                    }
                    umma_commit(&empty[s]);
                    if (in_g == kFlushBlocks - 1 || kb == num_kb - 1) umma_commit(&tmem_full[set]);
                }
        }
    } else {
        const int quarter = warp & 3;
        const int64_t row = int64_t(m_tile) * kBM + quarter * 32 + lane;
        float sum[kBN];
        float smax = -INFINITY, ssum = 0.f;                  // online softmax pair of this thread's row
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll
            for (int i = 0; i < kBN; ++i) sum[i] = 0.f;
#pragma unroll 1
            for (int g = 0; g < ngroups; ++g) {
                const int G = t * ngroups + g, set = G & 1;
                mbar_wait_bounded(&tmem_full[set], (G >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool two = num_kb - g * kFlushBlocks >= 2;                         // the group used both accumulators
                const uint32_t t0 = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(set * 2) * kBN;
#pragma unroll
                for (int c = 0; c < kBN; c += 32) {
                    uint32_t v[32], w[32];
                    tmem_ld_32x32_issue(t0 + uint32_t(c), v);
                    if (two) tmem_ld_32x32_issue(t0 + uint32_t(kBN + c), w);
                    tmem_ld_wait();
                    if (two) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            sum[c + i] = __fadd_rn(sum[c + i], __fadd_rn(__uint_as_float(v[i]), __uint_as_float(w[i])));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[c + i] = __fadd_rn(sum[c + i], __uint_as_float(v[i]));
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[set]);
            }
            // the tile is complete in registers: store it (the MMAs of the next tile are already running)
            const int64_t col0 = int64_t(n_tile0 + t) * kBN;
            if (row < M) {
                float *dst = Cout + row * ldc + col0;
                const bool whole = col0 + kBN <= Nn;
                const bool vec_ok = (ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(Cout) % 16 == 0) && whole;
                if (vec_ok) {
#pragma unroll
                    for (int i = 0; i < kBN; i += 4) *reinterpret_cast<float4 *>(dst + i) = make_float4(sum[i], sum[i + 1], sum[i + 2], sum[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < kBN; ++i)
                        if (col0 + i < Nn) dst[i] = sum[i];
                }
                if (row_stats != nullptr) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < kBN; ++i) {
                        sum[i] = (whole || col0 + i < Nn) ? __fmul_rn(a, sum[i]) : -INFINITY;
                        mx = fmaxf(mx, sum[i]);
                    }
                    if (mx > smax) {
                        ssum *= (smax == -INFINITY) ? 0.f : expf(smax - mx);
                        smax = mx;
                    }
                    const float ms = (smax == -INFINITY) ? 0.f : smax;
#pragma unroll
                    for (int i = 0; i < kBN; ++i) ssum += expf(__fsub_rn(sum[i], ms));       // exp(-inf) = 0 for the padding
                }
            }
        }
        if (row_stats != nullptr && row < M) {
            row_stats[row * 2] = smax;
            row_stats[row * 2 + 1] = ssum;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// S = exp(a * P - max) / sum from the row pairs left by the GEMM epilogue: per element the reference's operation
// sequence (separately rounded a * P, expf, true division); warp per row, coalesced.
__global__ void __launch_bounds__(256)
softmax_from_stats_kernel(const float *__restrict__ P, int64_t ldp, const float *__restrict__ row_stats, int64_t n_rows,
                          int n_cols, float a, float *__restrict__ S, int64_t lds) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float rm = row_stats[row * 2], rs = row_stats[row * 2 + 1];
    const float *src = P + row * ldp;
    float *dst = S + row * lds;
    for (int c = lane; c < lds; c += 32) dst[c] = c < n_cols ? __fdiv_rn(expf(__fsub_rn(__fmul_rn(a, src[c]), rm)), rs) : 0.f;
}

// ---- host ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 encode_fn2() {
    static EncodeTiledFn2 fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn2>(p);
    }();
    return fn;
}

static bool make_operand_map(CUtensorMap *map, const float *X, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn2 enc = encode_fn2();
    if (!enc) return false;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(cols) * sizeof(float)};
    cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(X), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t sim_matrix_tc_workspace(int64_t N, int64_t C, int64_t D) {
    const int64_t Np = ceil_div<int64_t>(N, kBM) * kBM, Cp = ceil_div<int64_t>(C, kBN) * kBN, Dp = ceil_div<int64_t>(D, kBK) * kBK;
    return size_t(Np + Cp) * size_t(Dp) * 2 * sizeof(float) + size_t(Np) * 2 * sizeof(float) /*row softmax pairs*/ + 1024;
}

// returns MCD_ERR_UNSUPPORTED when the tensor-map encoder is unavailable (caller then uses the CUDA-core kernel).
// kind: 0 = the streaming kernel (a CTA walks all column tiles of its 128-row band); the caller runs the stand-alone softmax;
//       2 = one CTA per output tile (round 1's kernel), ditto;
//       3 = the band kernel that also rescales the band inside the GEMM kernel;
//       4 = the streaming kernel whose epilogue keeps the row softmax pairs + softmax_from_stats_kernel.
// *S_done = 1 when S has been produced.
int sim_matrix_tc(const float *I, int64_t ldi, const float *T, int64_t ldt, int64_t N, int64_t C, int64_t D,
                  int normalize_rows, float *P, int64_t ldp, float *S, int64_t lds, float a, void *ws, size_t ws_bytes,
                  int kind, int *S_done, cudaStream_t st) {
    const int64_t Np = ceil_div<int64_t>(N, kBM) * kBM, Cp = ceil_div<int64_t>(C, kBN) * kBN, Dp = ceil_div<int64_t>(D, kBK) * kBK;
    *S_done = 0;
    if (ws_bytes < sim_matrix_tc_workspace(N, C, D)) return MCD_ERR_WORKSPACE;
    if (Np / kBM > 65535) return MCD_ERR_UNSUPPORTED;
    char *w = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    float *Ihi = reinterpret_cast<float *>(w);
    float *Ilo = Ihi + Np * Dp;
    float *Thi = Ilo + Np * Dp;
    float *Tlo = Thi + Cp * Dp;
    float *row_stats = Tlo + Cp * Dp;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    if (!make_operand_map(&mAhi, Ihi, Np, Dp, kBM) || !make_operand_map(&mAlo, Ilo, Np, Dp, kBM) ||
        !make_operand_map(&mBhi, Thi, Cp, Dp, kBN) || !make_operand_map(&mBlo, Tlo, Cp, Dp, kBN))
        return MCD_ERR_UNSUPPORTED;
    launch_pdl((prepare_rows_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(Np, 8))), dim3(256), 0, st, I, ldi, N, D, normalize_rows, Ihi, Ilo, Np, Dp);
    launch_pdl((prepare_rows_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(Cp, 8))), dim3(256), 0, st, T, ldt, C, D, normalize_rows, Thi, Tlo, Cp, Dp);
    count_launch(2);
    const int n_tiles = static_cast<int>(Cp / kBN), num_kb = static_cast<int>(Dp / kBK);
    if (kind == 3 && S != nullptr) {
        if (cudaFuncSetAttribute(gemm_tf32x3_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTcSmemBytes)) != cudaSuccess)
            return MCD_ERR_CUDA;
        launch_pdl((gemm_tf32x3_band_kernel), dim3(static_cast<unsigned>(Np / kBM)), dim3(kTcThreads), kTcSmemBytes, st, mAhi, mAlo, mBhi, mBlo, N, C, num_kb, n_tiles, P, ldp, S, lds, a);
        *S_done = 1;
        return check_launch();
    }
    if (kind == 2 || kind == 3) {
        if (cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTcSmemBytes)) != cudaSuccess)
            return MCD_ERR_CUDA;
        dim3 grid(static_cast<unsigned>(n_tiles), static_cast<unsigned>(Np / kBM));
        launch_pdl((gemm_tf32x3_kernel), dim3(grid), dim3(kTcThreads), kTcSmemBytes, st, mAhi, mAlo, mBhi, mBlo, N, C, num_kb, P, ldp);
        return check_launch();
    }
    if (cudaFuncSetAttribute(gemm_tf32x3_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTcSmemBytes)) != cudaSuccess)
        return MCD_ERR_CUDA;
    const int64_t tpc = tunable(kGemmTilesPerCta);
    const bool stats = kind == 4 && S != nullptr;
    // with the row pairs the whole band must pass through one CTA; without, any tile count per CTA will do
    // (default: half a band per CTA -- 3 tiles for the 763-concept set; measured at N = 100k: 1 tile 0.80 ms, 2: 0.75,
    // 3: 0.73, 6: 0.77: fewer, longer CTAs save pipeline fills, more of them balance the last wave)
    int tiles_per_cta = n_tiles >= 4 ? (n_tiles + 1) / 2 : n_tiles;
    if (stats) tiles_per_cta = n_tiles;
    else if (tpc > 0 && tpc <= n_tiles) tiles_per_cta = static_cast<int>(tpc);
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(n_tiles, tiles_per_cta)), static_cast<unsigned>(Np / kBM), 1);
    launch_pdl((gemm_tf32x3_long_kernel), dim3(grid), dim3(kTcThreads), kTcSmemBytes, st, mAhi, mAlo, mBhi, mBlo, N, C, num_kb, num_kb, P, ldp, 0,
                                                                    tiles_per_cta, n_tiles, a, stats ? row_stats : nullptr,
                                                                    tunable(kGemmDebugTerms) == 1 ? 1 : 3);
    int rc = check_launch();
    if (rc != MCD_OK || !stats) return rc;
    launch_pdl((softmax_from_stats_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(N, 8))), dim3(256), 0, st, P, ldp, row_stats, N, int(C), a, S, lds);
    *S_done = 1;
    return check_launch();
}

// ---- cos / cos^3 host: column statistics, transposed hi/lo operands (neuron slabs bound the workspace), long GEMM ---------
constexpr int64_t kCosSlab = 8192;                       // neurons per slab
constexpr int kMaxStatChunks = 64;

struct CosLayout {
    int64_t Np, Cp, Kp;
    int chunks, splits, kb_per_split;
    int64_t rows_per_chunk;
    size_t stat_off, p_off, a_off, part_off, total;
};
static CosLayout cos_layout(int64_t N, int64_t K, int64_t C) {
    CosLayout l;
    l.Np = ceil_div<int64_t>(N, kBK) * kBK;
    l.Cp = ceil_div<int64_t>(C, kBN) * kBN;
    const int64_t Ks = K < kCosSlab ? K : kCosSlab;
    l.Kp = ceil_div<int64_t>(Ks, kBM) * kBM;
    // row chunks of the statistics passes: enough blocks for the machine with the narrower matrix
    int64_t ch = ceil_div<int64_t>(4 * int64_t(num_sms()), ceil_div<int64_t>(C < K ? C : K, 32));
    if (ch > kMaxStatChunks) ch = kMaxStatChunks;
    if (ch > N / 64) ch = N / 64;
    if (ch < 1) ch = 1;
    l.rows_per_chunk = ceil_div<int64_t>(N, ch);
    l.chunks = static_cast<int>(ceil_div<int64_t>(N, l.rows_per_chunk));
    // split-K when the output tiles alone leave SMs idle (a split keeps >= 16 k-blocks)
    const int64_t tiles = (l.Cp / kBN) * (l.Kp / kBM), nkb = l.Np / kBK;
    int64_t sp = int64_t(num_sms()) / tiles;              // one wave: tiles x splits <= SMs
    if (sp > nkb / 16) sp = nkb / 16;
    if (sp > 32) sp = 32;
    if (sp < 1) sp = 1;
    l.kb_per_split = static_cast<int>(ceil_div<int64_t>(nkb, sp));
    l.splits = static_cast<int>(ceil_div<int64_t>(nkb, l.kb_per_split));
    auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
    l.stat_off = 1024;                                                   // alignment slack in front
    l.p_off = l.stat_off + up(size_t(2) * l.chunks * size_t(C + K) * sizeof(float));
    l.a_off = l.p_off + up(size_t(l.Cp) * l.Np * 2 * sizeof(float));
    l.part_off = l.a_off + up(size_t(l.Kp) * l.Np * 2 * sizeof(float));
    l.total = l.part_off + (l.splits > 1 ? up(size_t(l.splits) * l.Kp * l.Cp * sizeof(float)) : 0);
    return l;
}

size_t cos_similarity_tc_workspace(int64_t N, int64_t K, int64_t C) { return cos_layout(N, K, C).total; }

template <int STAT>
static int launch_col_partial(const ColSrc &s0, const ColSrc &s1, int64_t N, const CosLayout &l, cudaStream_t st) {
    const int64_t widest = s0.M > s1.M ? s0.M : s1.M;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(widest, 32)), static_cast<unsigned>(l.chunks), 2);
    launch_pdl((col_partial_kernel<STAT>), grid, dim3(256), 0, st, s0, s1, N, l.rows_per_chunk);
    return check_launch();
}

// out [K, C] = f(A)^T f(P): the whole cos_similarity / cos_similarity_cubed call
int cos_similarity_tc(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K, int64_t C, int cubed,
                      float min_norm, float *out, int64_t ldo, void *ws, size_t ws_bytes, cudaStream_t st) {
    const CosLayout l = cos_layout(N, K, C);
    if (ws_bytes < l.total) return MCD_ERR_WORKSPACE;
    if (l.Np > (int64_t(1) << 31) - 64) return MCD_ERR_UNSUPPORTED;      // TMA coordinates are 32-bit
    if (!encode_fn2()) return MCD_ERR_UNSUPPORTED;
    char *w = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    float *sumP = reinterpret_cast<float *>(w + l.stat_off - 1024);
    float *sumA = sumP + size_t(l.chunks) * C;
    float *sqP = sumA + size_t(l.chunks) * K;
    float *sqA = sqP + size_t(l.chunks) * C;
    float *Phi = reinterpret_cast<float *>(w + l.p_off - 1024), *Plo = Phi + l.Cp * l.Np;
    float *Ahi = reinterpret_cast<float *>(w + l.a_off - 1024), *Alo = Ahi + l.Kp * l.Np;
    float *part = reinterpret_cast<float *>(w + l.part_off - 1024);
    int rc;
    // ---- column statistics of both matrices ----
    auto fold = [&](float *p0, float *p1) {
        dim3 fgrid(static_cast<unsigned>(ceil_div<int64_t>(C > K ? C : K, 256)), 2);
        launch_pdl((col_fold_kernel), fgrid, dim3(256), 0, st, p0, C, p1, K, l.chunks);
        return check_launch();
    };
    if (cubed) {
        if ((rc = launch_col_partial<0>(ColSrc{P, ldp, C, nullptr, sumP}, ColSrc{A, lda, K, nullptr, sumA}, N, l, st)) != MCD_OK) return rc;
        if ((rc = fold(sumP, sumA)) != MCD_OK) return rc;
        if ((rc = launch_col_partial<2>(ColSrc{P, ldp, C, sumP, sqP}, ColSrc{A, lda, K, sumA, sqA}, N, l, st)) != MCD_OK) return rc;
    } else {
        if ((rc = launch_col_partial<1>(ColSrc{P, ldp, C, nullptr, sqP}, ColSrc{A, lda, K, nullptr, sqA}, N, l, st)) != MCD_OK) return rc;
    }
    if ((rc = fold(sqP, sqA)) != MCD_OK) return rc;
    const int folded = 1;                        // consumers read one row of the (folded) partial arrays
    if (cudaFuncSetAttribute(gemm_tf32x3_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTcSmemBytes)) != cudaSuccess)
        return MCD_ERR_CUDA;
    CUtensorMap mBhi, mBlo;
    if (!make_operand_map(&mBhi, Phi, l.Cp, l.Np, kBN) || !make_operand_map(&mBlo, Plo, l.Cp, l.Np, kBN)) return MCD_ERR_UNSUPPORTED;
    dim3 pgrid(static_cast<unsigned>(l.Cp / 32), static_cast<unsigned>(l.Np / 32));
    if (cubed) launch_pdl((prepare_cols_kernel<kPrepCos3>), pgrid, dim3(256), 0, st, P, ldp, N, C, sumP, sqP, folded, C, min_norm, Phi, Plo, l.Np);
    else launch_pdl((prepare_cols_kernel<kPrepCos>), pgrid, dim3(256), 0, st, P, ldp, N, C, static_cast<const float *>(nullptr), sqP, folded, C, min_norm, Phi, Plo, l.Np);
    if ((rc = check_launch()) != MCD_OK) return rc;
    const int total_kb = static_cast<int>(l.Np / kBK);
    for (int64_t k0 = 0; k0 < K; k0 += kCosSlab) {
        const int64_t kn = K - k0 < kCosSlab ? K - k0 : kCosSlab, kp = ceil_div<int64_t>(kn, kBM) * kBM;
        CUtensorMap mAhi, mAlo;
        if (!make_operand_map(&mAhi, Ahi, kp, l.Np, kBM) || !make_operand_map(&mAlo, Alo, kp, l.Np, kBM)) return MCD_ERR_UNSUPPORTED;
        dim3 agrid(static_cast<unsigned>(kp / 32), static_cast<unsigned>(l.Np / 32));
        // (the partial arrays are indexed by the column inside the whole matrix: offset pointer, full width K)
        if (cubed)
            launch_pdl((prepare_cols_kernel<kPrepCos3>), agrid, dim3(256), 0, st, A + k0, lda, N, kn, sumA + k0, sqA + k0, folded, K, min_norm, Ahi, Alo, l.Np);
        else
            launch_pdl((prepare_cols_kernel<kPrepCos>), agrid, dim3(256), 0, st, A + k0, lda, N, kn, static_cast<const float *>(nullptr), sqA + k0, folded, K, min_norm, Ahi, Alo, l.Np);
        if ((rc = check_launch()) != MCD_OK) return rc;
        dim3 grid(static_cast<unsigned>(l.Cp / kBN), static_cast<unsigned>(kp / kBM), static_cast<unsigned>(l.splits));
        if (l.splits == 1) {
            launch_pdl((gemm_tf32x3_long_kernel), dim3(grid), dim3(kTcThreads), kTcSmemBytes, st, mAhi, mAlo, mBhi, mBlo, kn, C, total_kb, l.kb_per_split,
                                                                            out + k0 * ldo, ldo, 0, 1, int(l.Cp / kBN), 1.f, nullptr, 3);
            if ((rc = check_launch()) != MCD_OK) return rc;
        } else {
            launch_pdl((gemm_tf32x3_long_kernel), dim3(grid), dim3(kTcThreads), kTcSmemBytes, st, mAhi, mAlo, mBhi, mBlo, kn, C, total_kb, l.kb_per_split,
                                                                            part, l.Cp, l.Kp * l.Cp, 1, int(l.Cp / kBN), 1.f, nullptr, 3);
            if ((rc = check_launch()) != MCD_OK) return rc;
            dim3 cgrid(static_cast<unsigned>(ceil_div<int64_t>(C, 256)), static_cast<unsigned>(kn));
            launch_pdl((splitk_combine_kernel), cgrid, dim3(256), 0, st, part, l.splits, kn, C, l.Cp, l.Kp * l.Cp, out + k0 * ldo, ldo);
            if ((rc = check_launch()) != MCD_OK) return rc;
        }
    }
    return MCD_OK;
}

}  // namespace mcd
