// K4 -- spatial pooling of a hooked NCHW activation: [B,C,H,W] -> [B,C]
// (replaces output.mean(dim=[2,3]) / output.amax(dim=[2,3]) inside the forward hook,
//  concept_vit/utils.py:38 and :47).
//
// Pure bandwidth: every activation element is read exactly once with 16-byte loads and reduced
// in fp32 (sum: 4-8 independent accumulators per thread, then warp/CTA trees; max: NaN-propagating
// like torch.amax).  Large planes (EfficientNet-B5 stem blocks are 760x456 = 1.4 MB per channel)
// are split across several CTAs so that even a 4-image batch of a 24-channel layer fills 148 SMs;
// the per-split partials are combined in split order by a second tiny kernel, so the result does
// not depend on scheduling.  Small planes (48x29) get one warp per plane.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace mcd {

constexpr int kPoolThreads = 256;

template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kVec = 4;
    __device__ static float to_f(float v) { return v; }
    __device__ static float from_f(float v) { return v; }
    __device__ static void unpack(const uint4 &q, float *f) {
        f[0] = __uint_as_float(q.x); f[1] = __uint_as_float(q.y); f[2] = __uint_as_float(q.z); f[3] = __uint_as_float(q.w);
    }
};
template <> struct Elem<__half> {
    static constexpr int kVec = 8;
    __device__ static float to_f(__half v) { return __half2float(v); }
    __device__ static __half from_f(float v) { return __float2half_rn(v); }
    __device__ static void unpack(const uint4 &q, float *f) {
        const __half2 *h = reinterpret_cast<const __half2 *>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
    __device__ static void unpack(const uint4 &q, float *f) {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
};

__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <bool MAX>
__device__ __forceinline__ float red_op(float acc, float v) {
    if (MAX) return (v > acc || v != v) ? v : acc;   // NaN sticks, as in torch.amax
    return acc + v;
}
template <bool MAX>
__device__ __forceinline__ float red_identity() { return MAX ? -INFINITY : 0.f; }

// Reduce elements [beg, end) of `x` with `nthreads` cooperating threads (thread `t` of them).
template <typename T, bool MAX>
__device__ __forceinline__ float reduce_range(const T *x, int64_t beg, int64_t end, int t, int nthreads) {
    constexpr int V = Elem<T>::kVec;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = red_identity<MAX>();
    // scalar head up to a 16-byte boundary
    const uintptr_t addr = reinterpret_cast<uintptr_t>(x + beg);
    int64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
    if (head > end - beg) head = end - beg;
    for (int64_t i = beg + t; i < beg + head; i += nthreads) acc[0] = red_op<MAX>(acc[0], Elem<T>::to_f(x[i]));
    const int64_t vbeg = beg + head;
    const int64_t nvec = (end - vbeg) / V;
    const uint4 *xv = reinterpret_cast<const uint4 *>(x + vbeg);
    int64_t i = t;
    for (; i + nthreads < nvec; i += 2 * nthreads) {       // two loads in flight per thread
        const uint4 q0 = ldg_stream_u4(xv + i);
        const uint4 q1 = ldg_stream_u4(xv + i + nthreads);
        float f0[8], f1[8];
        Elem<T>::unpack(q0, f0);
        Elem<T>::unpack(q1, f1);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = red_op<MAX>(red_op<MAX>(acc[e], f0[e]), f1[e]);
    }
    for (; i < nvec; i += nthreads) {
        const uint4 q0 = ldg_stream_u4(xv + i);
        float f0[8];
        Elem<T>::unpack(q0, f0);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = red_op<MAX>(acc[e], f0[e]);
    }
    for (int64_t j = vbeg + nvec * V + t; j < end; j += nthreads) acc[0] = red_op<MAX>(acc[0], Elem<T>::to_f(x[j]));
    float r = acc[0];
#pragma unroll
    for (int e = 1; e < V; ++e) r = red_op<MAX>(r, acc[e]);
    return r;
}

template <bool MAX>
__device__ __forceinline__ float warp_red(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<MAX>(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one warp per plane
template <typename T, bool MAX>
__global__ void __launch_bounds__(kPoolThreads)
pool_small_kernel(const T *__restrict__ x, int64_t planes, int64_t hw, T *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t plane = int64_t(blockIdx.x) * (kPoolThreads / 32) + (threadIdx.x >> 5);
    if (plane >= planes) return;
    float r = warp_red<MAX>(reduce_range<T, MAX>(x, plane * hw, (plane + 1) * hw, lane, 32));
    if (lane == 0) out[plane] = Elem<T>::from_f(MAX ? r : r / static_cast<float>(hw));
}

// one CTA per (plane, split); splits == 1 writes the result, otherwise a partial
template <typename T, bool MAX>
__global__ void __launch_bounds__(kPoolThreads)
pool_large_kernel(const T *__restrict__ x, int64_t hw, int splits, int64_t chunk, T *__restrict__ out,
                  float *__restrict__ partials) {
    __shared__ float s_red[kPoolThreads / 32];
    const int64_t plane = blockIdx.x / splits;
    const int split = static_cast<int>(blockIdx.x - plane * splits);
    const int64_t beg = plane * hw + int64_t(split) * chunk;
    const int64_t end = plane * hw + min(hw, int64_t(split + 1) * chunk);
    float r = warp_red<MAX>(reduce_range<T, MAX>(x, beg, end, threadIdx.x, kPoolThreads));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = s_red[0];
#pragma unroll
        for (int w = 1; w < kPoolThreads / 32; ++w) t = red_op<MAX>(t, s_red[w]);
        if (splits == 1) out[plane] = Elem<T>::from_f(MAX ? t : t / static_cast<float>(hw));
        else partials[blockIdx.x] = t;
    }
}

template <typename T, bool MAX>
__global__ void __launch_bounds__(kPoolThreads)
pool_finish_kernel(const float *__restrict__ partials, int64_t planes, int splits, int64_t hw, T *__restrict__ out) {
    const int64_t plane = int64_t(blockIdx.x) * kPoolThreads + threadIdx.x;
    if (plane >= planes) return;
    float t = partials[plane * splits];
    for (int s = 1; s < splits; ++s) t = red_op<MAX>(t, partials[plane * splits + s]);
    out[plane] = Elem<T>::from_f(MAX ? t : t / static_cast<float>(hw));
}

static int pool_splits(int64_t planes, int64_t hw) {
    if (hw <= 4096) return 0;   // warp-per-plane kernel
    const int64_t target = 4 * int64_t(num_sms());
    int64_t s = ceil_div<int64_t>(target, planes);
    const int64_t max_s = hw / 8192 > 1 ? hw / 8192 : 1;   // at least 8192 elements per CTA
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return static_cast<int>(s);
}

template <typename T, bool MAX>
static int pool_launch(const void *xv, int64_t planes, int64_t hw, void *outv, void *ws, size_t ws_bytes,
                       cudaStream_t st) {
    const T *x = static_cast<const T *>(xv);
    T *out = static_cast<T *>(outv);
    const int splits = pool_splits(planes, hw);
    if (splits == 0) {
        const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(planes, kPoolThreads / 32));
        pool_small_kernel<T, MAX><<<grid, kPoolThreads, 0, st>>>(x, planes, hw, out);
        return check_launch();
    }
    if (planes * splits > 0x7FFFFFFFll) return MCD_ERR_UNSUPPORTED;
    if (splits > 1 && (!ws || ws_bytes < size_t(planes) * splits * sizeof(float))) return MCD_ERR_WORKSPACE;
    int64_t chunk = ceil_div<int64_t>(hw, splits);
    chunk = ceil_div<int64_t>(chunk, 32) * 32;   // keep split starts 128-byte friendly
    pool_large_kernel<T, MAX><<<static_cast<unsigned>(planes * splits), kPoolThreads, 0, st>>>(
        x, hw, splits, chunk, out, static_cast<float *>(ws));
    int rc = check_launch();
    if (rc != MCD_OK || splits == 1) return rc;
    pool_finish_kernel<T, MAX><<<static_cast<unsigned>(ceil_div<int64_t>(planes, kPoolThreads)), kPoolThreads, 0, st>>>(
        static_cast<const float *>(ws), planes, splits, hw, out);
    return check_launch();
}

}  // namespace mcd

extern "C" size_t mcd_pool_nchw_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    const int s = mcd::pool_splits(B * C, H * W);
    return s > 1 ? size_t(B * C) * s * sizeof(float) : 0;
}

extern "C" int mcd_pool_nchw(const void *x, mcd_dtype_t dtype, int64_t B, int64_t C, int64_t H, int64_t W,
                             mcd_pool_t mode, void *out, void *workspace, size_t workspace_bytes,
                             mcd_stream_t stream) {
    using namespace mcd;
    if (!x || !out || B < 1 || C < 1 || H < 1 || W < 1) return MCD_ERR_INVALID_ARGUMENT;
    if (mode != MCD_POOL_MEAN && mode != MCD_POOL_MAX) return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t planes = B * C, hw = H * W;
    const bool mx = mode == MCD_POOL_MAX;
    switch (dtype) {
        case MCD_F32:
            return mx ? pool_launch<float, true>(x, planes, hw, out, workspace, workspace_bytes, st)
                      : pool_launch<float, false>(x, planes, hw, out, workspace, workspace_bytes, st);
        case MCD_F16:
            return mx ? pool_launch<__half, true>(x, planes, hw, out, workspace, workspace_bytes, st)
                      : pool_launch<__half, false>(x, planes, hw, out, workspace, workspace_bytes, st);
        case MCD_BF16:
            return mx ? pool_launch<__nv_bfloat16, true>(x, planes, hw, out, workspace, workspace_bytes, st)
                      : pool_launch<__nv_bfloat16, false>(x, planes, hw, out, workspace, workspace_bytes, st);
        default:
            return MCD_ERR_INVALID_ARGUMENT;
    }
}
