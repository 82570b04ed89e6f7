// Shared helpers for the sm_100a kernels behind include/mcd_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mcd_b200.h"

namespace mcd {

constexpr int kNumSMsB200 = 148;

// ---- launch accounting / error plumbing -----------------------------------------------------
void count_launch(int n = 1);          // abi.cu
int64_t tunable(int which);            // abi.cu
enum Tunable { kTopkSplits = 0, kAccumTile = 1, kTopkVariant = 2, kAccumUnroll = 3, kTopkCols = 4, kTopkStages = 5, kTopkOcc = 6, kGemmVariant = 7,
               kTopkPre = 8, kTopkSmall = 9, kTopkFilter = 10, kFilterStages = 11, kFilterChunkTiles = 12, kPipeChunks = 13, kFilterOrder = 14, kAccumPadKb = 15,
               kGemmTilesPerCta = 16, kGemmDebugTerms = 17, kNumTunables = 18 };
int num_sms();                         // abi.cu (cached cudaDevAttrMultiProcessorCount)

int debug_sync_check();                // abi.cu: MCD_DEBUG_SYNC=1 -> synchronise after every launch and report the failing one

inline int check_launch(int n = 1) {
    count_launch(n);
    if (cudaPeekAtLastError() != cudaSuccess) return MCD_ERR_CUDA;
    return debug_sync_check();
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- ordered keys for the top-k total order ---------------------------------------------------
// value desc, index asc, NaN largest, -0.0 == +0.0.  key(v) is a u32 that sorts like v under
// that rule; 0 is reserved as the "empty slot" sentinel (no finite/inf/NaN value maps to it).
__device__ __forceinline__ uint32_t ordered_key(float v) {
    if (v != v) return 0xFFFFFFFFu;
    uint32_t b = __float_as_uint(v + 0.0f);             // -0.0 + 0.0 == +0.0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// A float t such that `!(v <= t)` holds for every v whose key exceeds `key` (and may hold for
// equal keys only when key is the sentinel / NaN, where t = NaN admits everything).
__device__ __forceinline__ float key_to_threshold(uint32_t key) {
    if (key == 0u || key == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
    uint32_t b = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
    return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long pack_key(uint32_t hi, uint32_t lo) {
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

// ---- programmatic dependent launch (the kernels of one call form a chain on one stream) ---------------------------------
// Every kernel of the chain starts with pdl_enter(): it waits until the grid in front of it has completed and its writes
// are visible (a no-op for a normally launched kernel), then lets the grid behind it be scheduled as soon as all CTAs of
// this one are resident or done.  The successor's CTAs are then already on the SMs, parked in their own pdl_enter(), when
// this grid's last CTAs retire: no launch latency and no empty-machine ramp between two dependent kernels (11 per call).
// EVERY CTA must pass pdl_enter() before it returns or touches global memory -- the guarantee is transitive only then.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);       // errors surface in check_launch()
}

// ---- PTX: mbarrier + bulk async copy (TMA engine, 1-D) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same on a shared::cta ADDRESS (hot loops keep the address in a register instead of re-deriving it from a
// generic pointer every iteration)
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_expect_tx_addr(uint32_t bar_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- thread-block clusters / distributed shared memory -----------------------------------------
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t smem_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_dsmem_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_dsmem_u32(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}

// ---- loads ------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_nc_v4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_nc_v4_hint(const float *p, uint64_t policy) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// bitonic sort, descending, of 32 * PER words held as word i = PER * lane + slot (strides below PER exchange inside a
// lane, the others with shfl.xor)
template <int PER, typename T>
__device__ __forceinline__ void bitonic_desc_regs(T (&v)[PER], int lane) {
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
                    const T o = __shfl_xor_sync(0xffffffffu, v[e], stride / PER);
                    const T hi = v[e] > o ? v[e] : o, lo = v[e] > o ? o : v[e];
                    v[e] = keep_max ? hi : lo;
                }
            } else {
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    if ((e & stride) == 0) {
                        const bool desc = size >= TOTAL || ((lane * PER + e) & size) == 0;
                        const T x = v[e], y = v[e + stride < PER ? e + stride : e];
                        const T hi = x > y ? x : y, lo = x > y ? y : x;
                        v[e] = desc ? hi : lo;
                        v[e + stride < PER ? e + stride : e] = desc ? lo : hi;
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mcd
