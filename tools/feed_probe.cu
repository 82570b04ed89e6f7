// Feed-structure probe for the K2 filter scan (not part of the product): the scan's exact feed -- single-warp CTAs, one
// work item = 128 columns x a row chunk, a private ring of [rows x 128] TMA tiles, lane 0 re-arms a stage as soon as the
// warp has pulled the tile into registers -- WITHOUT the filter work, so the number printed is what this feed structure
// can pull from HBM.  `pad` bytes of extra dynamic shared memory stand in for the append bags (they set the residency).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/feed_probe tools/feed_probe.cu
//   run:   build/feed_probe [N] [K]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}

template <int ROWS>
__global__ void __launch_bounds__(32)
feed_kernel(const __grid_constant__ CUtensorMap tmap, int64_t N, int chunk_tiles, int nstage, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x;
    const uint32_t ring = smem_u32(smem);
    constexpr uint32_t kTile = ROWS * 128 * 4;
    const uint32_t bars = ring + nstage * kTile;
    const int64_t tiles_total = (N + ROWS - 1) / ROWS;
    const int64_t t0 = (int64_t)blockIdx.y * chunk_tiles;
    const int ntiles = (int)(tiles_total - t0 < chunk_tiles ? tiles_total - t0 : chunk_tiles);
    if (ntiles <= 0) return;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (lane == 0) {
        for (int s = 0; s < nstage; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto arm = [&](int t, int s) {
        mbar_expect(bars + 8 * s, kTile);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                     ::"r"(ring + s * kTile), "l"(&tmap), "r"((int)(blockIdx.x * 128)), "r"((int)((t0 + t) * ROWS)), "r"(bars + 8 * s), "l"(pol) : "memory");
    };
    if (lane == 0)
        for (int s = 0; s < nstage && s < ntiles; ++s) arm(s, s);
    float acc = 0.f;
    int s = 0, use = 0;
    for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bars + 8 * s, use & 1);
        float4 v[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[r].x), "=f"(v[r].y), "=f"(v[r].z), "=f"(v[r].w) : "r"(ring + s * kTile + r * 512 + lane * 16));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && t + nstage < ntiles) arm(t + nstage, s);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc += v[r].x + v[r].y + v[r].z + v[r].w;
        if (++s == nstage) { s = 0; ++use; }
    }
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS>
static void run(EncodeTiledFn enc, float *A, int64_t N, int64_t K, int64_t lda, int nstage, int pad_kb, int items_per_slot, float *sink,
                cudaEvent_t e0, cudaEvent_t e1, int promo) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)lda * 4};
    cuuint32_t box[2] = {128, (cuuint32_t)ROWS};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return;
    }
    const size_t smem = (size_t)nstage * ROWS * 512 + 64 + (size_t)pad_kb * 1024;
    CK(cudaFuncSetAttribute(feed_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int resident = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, feed_kernel<ROWS>, 32, smem));
    const int64_t nblk = (K + 127) / 128, tiles_total = (N + ROWS - 1) / ROWS;
    int64_t ct = (tiles_total * nblk + (int64_t)148 * resident * items_per_slot - 1) / ((int64_t)148 * resident * items_per_slot);
    if (ct < 8) ct = 8;
    dim3 grid((unsigned)nblk, (unsigned)((tiles_total + ct - 1) / ct));
    float ms = 0;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        feed_kernel<ROWS><<<grid, 32, smem>>>(map, N, (int)ct, nstage, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    CK(cudaGetLastError());
    printf("rows %2d stages %d pad %2d KB promo %d lda %6lld: %2d CTAs/SM, %5.1f KB ring/SM, item %4lld tiles, grid (%u,%u)  %7.3f ms  %6.0f GB/s\n",
           ROWS, nstage, pad_kb, promo, (long long)lda, resident, resident * nstage * ROWS * 0.5, (long long)ct, grid.x, grid.y, ms,
           double(N) * K * 4 / 1e9 / ms * 1e3);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 100000, K = argc > 2 ? atoll(argv[2]) : 32768;
    float *A, *sink;
    const int64_t lda_pad = K + 64;
    CK(cudaMalloc(&A, N * lda_pad * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(A, 0, N * lda_pad * 4));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    // the product's configuration first: 8-row tiles, 2 stages, 10 KB of bags
    run<8>(enc, A, N, K, K, 2, 10, 32, sink, e0, e1, 1);
    run<8>(enc, A, N, K, K, 2, 10, 32, sink, e0, e1, 0);
    run<8>(enc, A, N, K, K, 2, 10, 32, sink, e0, e1, 2);
    run<8>(enc, A, N, K, lda_pad, 2, 10, 32, sink, e0, e1, 1);          // does the power-of-two row pitch matter?
    for (int pad : {0, 4, 6, 10})
        for (int st : {2, 3, 4, 6}) run<8>(enc, A, N, K, K, st, pad, 32, sink, e0, e1, 1);
    for (int pad : {0, 6, 10})
        for (int st : {2, 3, 4}) run<16>(enc, A, N, K, K, st, pad, 32, sink, e0, e1, 1);
    for (int pad : {0, 10})
        for (int st : {2, 3, 4}) run<32>(enc, A, N, K, K, st, pad, 32, sink, e0, e1, 1);
    for (int ips : {4, 8, 16, 64}) run<8>(enc, A, N, K, K, 2, 10, ips, sink, e0, e1, 1);
    return 0;
}
