// Bandwidth probe for the column-block access pattern of the top-k scan (not part of the product):
// how fast can a B200 stream A [N, K] fp32 when every CTA reads a `cols`-wide column block row by row?
//   mode tma : TMA tensor tiles [rows x cols] into a smem ring (producer lane + consumer warps that only wait/release)
//   mode ldg : 16-byte read-only loads straight to registers, U rows in flight per warp
//   mode flat: plain contiguous streaming read of the whole matrix (reference peak for this box)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/stream_probe tools/stream_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}

__global__ void __launch_bounds__(288, 1)
probe_tma(const __grid_constant__ CUtensorMap tmap, int64_t N, int cols, int rows, int nstage, int hint, int touch, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *ring = (float *)smem;
    uint64_t *full = (uint64_t *)(smem + (size_t)nstage * rows * cols * 4);
    uint64_t *empty = full + 16;
    const int tid = threadIdx.x;
    const int ntiles = (int)((N + rows - 1) / rows);
    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= 256) {
        if (tid == 256) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            int stage = 0, use = 0;
            for (int t = 0; t < ntiles; ++t) {
                if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1);
                mbar_expect(&full[stage], (uint32_t)(rows * cols * 4));
                void *dst = ring + (size_t)stage * rows * cols;
                if (hint)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                                 ::"r"(smem_u32(dst)), "l"(&tmap), "r"((int)(blockIdx.x * cols)), "r"(t * rows), "r"(smem_u32(&full[stage])), "l"(pol) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(dst)), "l"(&tmap), "r"((int)(blockIdx.x * cols)), "r"(t * rows), "r"(smem_u32(&full[stage])) : "memory");
                if (++stage == nstage) { stage = 0; ++use; }
            }
        }
        return;
    }
    const int warp = tid >> 5, lane = tid & 31;
    int stage = 0, use = 0;
    float acc = 0.f;
    for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&full[stage], use & 1);
        if (touch && lane * 4 < cols) {
            const float *tile = ring + (size_t)stage * rows * cols + lane * 4;
            for (int r = warp; r < rows; r += 8) {
                float4 v = *(const float4 *)(tile + r * cols);
                acc += v.x + v.y + v.z + v.w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == nstage) { stage = 0; ++use; }
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256) probe_ldg(const float *__restrict__ A, int64_t lda, int64_t N, int cols, float *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *base = A + (int64_t)blockIdx.x * cols + lane * 4;
    const bool ok = lane * 4 < cols;
    float acc = 0.f;
    for (int64_t r0 = (int64_t)warp * U; r0 < N; r0 += 8 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u] = make_float4(0, 0, 0, 0);
            if (ok && r0 + u < N)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(base + (r0 + u) * lda));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) probe_flat(const float4 *__restrict__ A, int64_t n4, float *sink) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldg(A + i), b = __ldg(A + i + stride), c = __ldg(A + i + 2 * stride), d = __ldg(A + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += stride) acc += __ldg(A + i).x;
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 100000, K = argc > 2 ? atoll(argv[2]) : 32768;
    float *A, *sink;
    CK(cudaMalloc(&A, N * K * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(A, 0, N * K * 4));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double gb = double(N) * K * 4 / 1e9;
    auto report = [&](const char *name, float ms) { printf("%-58s %8.3f ms  %7.0f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };

    {   // flat
        for (int it = 0; it < 2; ++it) {
            CK(cudaEventRecord(e0));
            probe_flat<<<148 * 8, 256>>>((const float4 *)A, N * K / 4, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        report("flat contiguous read", ms);
    }
    const int colsv[] = {32, 48};
    for (int cols : colsv) {
        if (cols <= 128) {
            char name[128];
            float ms;
            for (int it = 0; it < 2; ++it) { CK(cudaEventRecord(e0)); probe_ldg<8><<<(unsigned)((K + cols - 1) / cols), 256>>>(A, K, N, cols, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); }
            CK(cudaEventElapsedTime(&ms, e0, e1)); snprintf(name, sizeof name, "ldg cols=%d U=8", cols); report(name, ms);
            for (int it = 0; it < 2; ++it) { CK(cudaEventRecord(e0)); probe_ldg<16><<<(unsigned)((K + cols - 1) / cols), 256>>>(A, K, N, cols, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); }
            CK(cudaEventElapsedTime(&ms, e0, e1)); snprintf(name, sizeof name, "ldg cols=%d U=16", cols); report(name, ms);
        }
        const int rowsv[] = {16, 32, 64};
        for (int rows : rowsv) {
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
            cuuint64_t strides[1] = {(cuuint64_t)K * 4};
            cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)rows};
            cuuint32_t estr[2] = {1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("encode failed cols=%d rows=%d\n", cols, rows);
                continue;
            }
            const size_t tile = (size_t)rows * cols * 4;
            const int stv[] = {2, 4, 8};
            for (int nstage : stv) {
                if (nstage > 16) continue;
                const size_t smem = tile * nstage + 256;
                if (smem > 200 * 1024) continue;
                for (int hint = 0; hint < 2; ++hint) {
                    CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    float ms;
                    for (int it = 0; it < 2; ++it) {
                        CK(cudaEventRecord(e0));
                        probe_tma<<<(unsigned)((K + cols - 1) / cols), 288, smem>>>(map, N, cols, rows, nstage, hint, 1, sink);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                    }
                    CK(cudaGetLastError());
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    char name[128];
                    snprintf(name, sizeof name, "tma cols=%d rows=%d stages=%d (%zu KB in flight) hint=%d", cols, rows, nstage, tile * nstage / 1024, hint);
                    report(name, ms);
                }
            }
        }
    }
    return 0;
}
