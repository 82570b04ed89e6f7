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
//
// The result row of image b may live anywhere: out[b * out_ld + c], in the activation's dtype or in fp32 -- so the hook
// writes straight into the [n_images, sum K_l] fp32 activation matrix (hooks.ActivationStack) without an intermediate
// [B, C] tile -- and channels-last activations (memory order [B, H, W, C]) are pooled in place by pool_nhwc_kernel
// (threads along C, rows of H*W split over the CTA and, for large planes, over several CTAs) instead of being
// repacked to NCHW first.
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

// where the pooled value of plane (b, c) goes: out[b * out_ld + c], as TO (the activation's dtype or fp32)
template <typename TO>
struct OutRef {
    TO *out;
    int64_t C, out_ld;
    __device__ __forceinline__ void put(int64_t plane, float v) const {
        const int64_t b = plane / C;
        out[b * out_ld + (plane - b * C)] = Elem<TO>::from_f(v);
    }
};

// one warp per plane
template <typename T, typename TO, bool MAX>
__global__ void __launch_bounds__(kPoolThreads)
pool_small_kernel(const T *__restrict__ x, int64_t planes, int64_t hw, OutRef<TO> out) {
    const int lane = threadIdx.x & 31;
    const int64_t plane = int64_t(blockIdx.x) * (kPoolThreads / 32) + (threadIdx.x >> 5);
    if (plane >= planes) return;
    float r = warp_red<MAX>(reduce_range<T, MAX>(x, plane * hw, (plane + 1) * hw, lane, 32));
    if (lane == 0) out.put(plane, MAX ? r : r / static_cast<float>(hw));
}

// one CTA per (plane, split); splits == 1 writes the result, otherwise a partial
template <typename T, typename TO, bool MAX>
__global__ void __launch_bounds__(kPoolThreads)
pool_large_kernel(const T *__restrict__ x, int64_t hw, int splits, int64_t chunk, OutRef<TO> out,
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
        if (splits == 1) out.put(plane, MAX ? t : t / static_cast<float>(hw));
        else partials[blockIdx.x] = t;
    }
}

template <typename TO, bool MAX>
__global__ void __launch_bounds__(kPoolThreads)
pool_finish_kernel(const float *__restrict__ partials, int64_t planes, int splits, int64_t hw, OutRef<TO> out) {
    const int64_t plane = int64_t(blockIdx.x) * kPoolThreads + threadIdx.x;
    if (plane >= planes) return;
    float t = partials[plane * splits];
    for (int s = 1; s < splits; ++s) t = red_op<MAX>(t, partials[plane * splits + s]);
    out.put(plane, MAX ? t : t / static_cast<float>(hw));
}

// Channels-last activations: x is [B, HW, C] in memory.  blockDim = (TX, TY): thread (tx, ty) owns the V adjacent channels
// (blockIdx.x * TX + tx) * V ... and the rows ty, ty + TY, ... of its split of the H*W axis (a warp reads contiguous
// memory: consecutive rows of a channel tile follow each other when the tile spans all of C); the TY partial results
// are folded in ty order through shared memory.  grid = (channel tiles, B, splits); splits > 1 writes partials
// [plane][split] for pool_finish_kernel.
template <typename T, typename TO, bool MAX, int V>
__global__ void __launch_bounds__(kPoolThreads)
pool_nhwc_kernel(const T *__restrict__ x, int64_t C, int64_t hw, int splits, int64_t chunk, OutRef<TO> out,
                 float *__restrict__ partials) {
    __shared__ float s_red[kPoolThreads][V > 1 ? V : 1];
    const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    const int64_t c0 = (int64_t(blockIdx.x) * TX + tx) * V;
    const int64_t b = blockIdx.y;
    const int split = blockIdx.z;
    const int64_t r_beg = int64_t(split) * chunk, r_end = min(hw, r_beg + chunk);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = red_identity<MAX>();
    if (c0 < C) {
        const T *base = x + (b * hw) * C + c0;
        for (int64_t r = r_beg + ty; r < r_end; r += TY) {
            if (V > 1) {
                const uint4 q = ldg_stream_u4(base + r * C);
                float f[8];
                Elem<T>::unpack(q, f);
#pragma unroll
                for (int e = 0; e < V; ++e) acc[e] = red_op<MAX>(acc[e], f[e]);
            } else {
                acc[0] = red_op<MAX>(acc[0], Elem<T>::to_f(base[r * C]));
            }
        }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) s_red[ty * TX + tx][e] = acc[e];
    __syncthreads();
    if (ty == 0 && c0 < C) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (c0 + e < C) {
                float t = s_red[tx][e];
                for (int y = 1; y < TY; ++y) t = red_op<MAX>(t, s_red[y * TX + tx][e]);
                const int64_t plane = b * C + c0 + e;
                if (splits == 1) out.put(plane, MAX ? t : t / static_cast<float>(hw));
                else partials[plane * splits + split] = t;
            }
        }
    }
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

template <typename T, typename TO, bool MAX>
static int pool_launch(const void *xv, int64_t B, int64_t C, int64_t hw, void *outv, int64_t out_ld, void *ws,
                       size_t ws_bytes, cudaStream_t st) {
    const T *x = static_cast<const T *>(xv);
    const int64_t planes = B * C;
    OutRef<TO> out{static_cast<TO *>(outv), C, out_ld};
    const int splits = pool_splits(planes, hw);
    if (splits == 0) {
        const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(planes, kPoolThreads / 32));
        pool_small_kernel<T, TO, MAX><<<grid, kPoolThreads, 0, st>>>(x, planes, hw, out);
        return check_launch();
    }
    if (planes * splits > 0x7FFFFFFFll) return MCD_ERR_UNSUPPORTED;
    if (splits > 1 && (!ws || ws_bytes < size_t(planes) * splits * sizeof(float))) return MCD_ERR_WORKSPACE;
    int64_t chunk = ceil_div<int64_t>(hw, splits);
    chunk = ceil_div<int64_t>(chunk, 32) * 32;   // keep split starts 128-byte friendly
    pool_large_kernel<T, TO, MAX><<<static_cast<unsigned>(planes * splits), kPoolThreads, 0, st>>>(
        x, hw, splits, chunk, out, static_cast<float *>(ws));
    int rc = check_launch();
    if (rc != MCD_OK || splits == 1) return rc;
    pool_finish_kernel<TO, MAX><<<static_cast<unsigned>(ceil_div<int64_t>(planes, kPoolThreads)), kPoolThreads, 0, st>>>(
        static_cast<const float *>(ws), planes, splits, hw, out);
    return check_launch();
}

// channels-last: splits of the H*W axis so that (channel tiles x B x splits) CTAs cover the machine
static int nhwc_splits(int64_t B, int64_t C, int64_t hw, int tiles) {
    const int64_t target = 4 * int64_t(num_sms());
    int64_t s = ceil_div<int64_t>(target, B * tiles);
    const int64_t max_s = hw / 512 > 1 ? hw / 512 : 1;       // at least 512 rows per CTA
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    (void)C;
    return static_cast<int>(s);
}
struct NhwcShape {
    int V, TX, TY, tiles, splits;
};
template <typename T>
static NhwcShape nhwc_shape(const void *x, int64_t B, int64_t C, int64_t hw) {
    NhwcShape s;
    const bool vec = (C * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
    s.V = vec ? Elem<T>::kVec : 1;
    const int64_t nvec = ceil_div<int64_t>(C, s.V);
    s.TX = 1;
    while (s.TX < 32 && s.TX < nvec) s.TX <<= 1;
    s.TY = kPoolThreads / s.TX;
    s.tiles = static_cast<int>(ceil_div<int64_t>(nvec, s.TX));
    s.splits = nhwc_splits(B, C, hw, s.tiles);
    return s;
}

template <typename T, typename TO, bool MAX>
static int pool_launch_nhwc(const void *xv, int64_t B, int64_t C, int64_t hw, void *outv, int64_t out_ld, void *ws,
                            size_t ws_bytes, cudaStream_t st) {
    const T *x = static_cast<const T *>(xv);
    OutRef<TO> out{static_cast<TO *>(outv), C, out_ld};
    const NhwcShape sh = nhwc_shape<T>(xv, B, C, hw);
    if (B > 65535) return MCD_ERR_UNSUPPORTED;
    if (sh.splits > 1 && (!ws || ws_bytes < size_t(B * C) * sh.splits * sizeof(float))) return MCD_ERR_WORKSPACE;
    const int64_t chunk = ceil_div<int64_t>(hw, sh.splits);
    dim3 grid(static_cast<unsigned>(sh.tiles), static_cast<unsigned>(B), static_cast<unsigned>(sh.splits)), block(sh.TX, sh.TY);
    float *part = static_cast<float *>(ws);
    if (sh.V == 1) pool_nhwc_kernel<T, TO, MAX, 1><<<grid, block, 0, st>>>(x, C, hw, sh.splits, chunk, out, part);
    else pool_nhwc_kernel<T, TO, MAX, Elem<T>::kVec><<<grid, block, 0, st>>>(x, C, hw, sh.splits, chunk, out, part);
    int rc = check_launch();
    if (rc != MCD_OK || sh.splits == 1) return rc;
    pool_finish_kernel<TO, MAX><<<static_cast<unsigned>(ceil_div<int64_t>(B * C, kPoolThreads)), kPoolThreads, 0, st>>>(
        part, B * C, sh.splits, hw, out);
    return check_launch();
}

template <typename T, typename TO>
static int pool_dispatch(const void *x, int64_t B, int64_t C, int64_t hw, int channels_last, bool mx, void *out,
                         int64_t out_ld, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (channels_last)
        return mx ? pool_launch_nhwc<T, TO, true>(x, B, C, hw, out, out_ld, ws, ws_bytes, st)
                  : pool_launch_nhwc<T, TO, false>(x, B, C, hw, out, out_ld, ws, ws_bytes, st);
    return mx ? pool_launch<T, TO, true>(x, B, C, hw, out, out_ld, ws, ws_bytes, st)
              : pool_launch<T, TO, false>(x, B, C, hw, out, out_ld, ws, ws_bytes, st);
}

}  // namespace mcd

extern "C" size_t mcd_pool_nchw_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    // enough for either memory order: the channels-last kernel splits H*W the most when its channel tiles are the fewest
    // (8 channels per 16-byte load)
    const int s = mcd::pool_splits(B * C, H * W);
    const size_t nchw = s > 1 ? size_t(B * C) * s * sizeof(float) : 0;
    const int tiles_min = static_cast<int>(mcd::ceil_div<int64_t>(mcd::ceil_div<int64_t>(C, 8), 32));
    const int sl = mcd::nhwc_splits(B, C, H * W, tiles_min);
    const size_t nhwc = sl > 1 ? size_t(B * C) * sl * sizeof(float) : 0;
    return nchw > nhwc ? nchw : nhwc;
}

extern "C" int mcd_pool_nchw_to(const void *x, mcd_dtype_t dtype, int64_t B, int64_t C, int64_t H, int64_t W,
                                int channels_last, mcd_pool_t mode, void *out, mcd_dtype_t out_dtype, int64_t out_ld,
                                void *workspace, size_t workspace_bytes, mcd_stream_t stream) {
    using namespace mcd;
    if (!x || !out || B < 1 || C < 1 || H < 1 || W < 1 || out_ld < C) return MCD_ERR_INVALID_ARGUMENT;
    if (mode != MCD_POOL_MEAN && mode != MCD_POOL_MAX) return MCD_ERR_INVALID_ARGUMENT;
    if (out_dtype != dtype && out_dtype != MCD_F32) return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t hw = H * W;
    const bool mx = mode == MCD_POOL_MAX;
    const bool f32out = out_dtype == MCD_F32;
    switch (dtype) {
        case MCD_F32:
            return pool_dispatch<float, float>(x, B, C, hw, channels_last, mx, out, out_ld, workspace, workspace_bytes, st);
        case MCD_F16:
            return f32out ? pool_dispatch<__half, float>(x, B, C, hw, channels_last, mx, out, out_ld, workspace, workspace_bytes, st)
                          : pool_dispatch<__half, __half>(x, B, C, hw, channels_last, mx, out, out_ld, workspace, workspace_bytes, st);
        case MCD_BF16:
            return f32out ? pool_dispatch<__nv_bfloat16, float>(x, B, C, hw, channels_last, mx, out, out_ld, workspace, workspace_bytes, st)
                          : pool_dispatch<__nv_bfloat16, __nv_bfloat16>(x, B, C, hw, channels_last, mx, out, out_ld, workspace,
                                                                        workspace_bytes, st);
        default:
            return MCD_ERR_INVALID_ARGUMENT;
    }
}

extern "C" int mcd_pool_nchw(const void *x, mcd_dtype_t dtype, int64_t B, int64_t C, int64_t H, int64_t W,
                             mcd_pool_t mode, void *out, void *workspace, size_t workspace_bytes,
                             mcd_stream_t stream) {
    return mcd_pool_nchw_to(x, dtype, B, C, H, W, 0, mode, out, dtype, C, workspace, workspace_bytes, stream);
}
