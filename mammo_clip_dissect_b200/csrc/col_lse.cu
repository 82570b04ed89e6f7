// K3b -- log p(d) over the neurons of a call and the final subtraction
// (replaces torch.logsumexp(L, dim=0) - log(K) and `L - lam * prob_d`,
//  concept_vit/similarity.py:67-72 and :91-96).
//
// The log-sum-exp over neurons is the one place where neurons are coupled, so it is also the one
// exchange point of the neuron-sharded multi-GPU path.  To make results independent of how the
// neurons are sharded it is computed in fixed blocks of MCD_LSE_BLOCK = 256 neurons:
//   partial[b] = ( m_b[c] = max_j L[j,c],  s_b[c] = sum_j exp(L[j,c] - m_b[c]) )   (fp32, j in block order)
// and the partials of ALL blocks are combined in global block order in fp64:
//   lse[c] = M + log( sum_b s_b * exp(m_b - M) ),  M = max_b m_b   (fp64 sum, fp32 block weights).
#include "common.cuh"

namespace mcd {

constexpr int kLseThreads = 128;

// thread per concept, one pass over the block's 256 rows with an online (max, sum) pair: the running sum is
// rescaled whenever the maximum rises (the rescale factor is exactly 1 otherwise), 16 rows per step
__global__ void __launch_bounds__(kLseThreads)
col_lse_partials_kernel(const float *__restrict__ L, int64_t ldl, int64_t K, int C, float *__restrict__ partials,
                        const int32_t *__restrict__ block_tab) {
    pdl_enter();
    const int c = blockIdx.x * kLseThreads + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (c >= C) return;
    // one call: blocks of 256 neurons from row 0; several layers in one matrix: the table lists every block's rows
    const int64_t j0 = block_tab ? block_tab[b * 3 + 0] : b * MCD_LSE_BLOCK;
    const int64_t j1 = block_tab ? j0 + block_tab[b * 3 + 1] : min(K, j0 + MCD_LSE_BLOCK);
    const float *col = L + c;
    float m = -INFINITY, s = 0.f;
    int64_t j = j0;
    constexpr int kAhead = 16;      // independent row loads in flight per thread (20 warps per SM x 16 x 128 B)
    for (; j + kAhead <= j1; j += kAhead) {
        float x[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) x[u] = col[(j + u) * ldl];
        float mx = x[0];
#pragma unroll
        for (int u = 1; u < kAhead; ++u) mx = fmaxf(mx, x[u]);
        if (mx > m) {
            s *= (m == -INFINITY) ? 0.f : expf(m - mx);
            m = mx;
        }
        const float ms = (m == -INFINITY) ? 0.f : m;
#pragma unroll
        for (int u = 0; u < kAhead; ++u) s += expf(x[u] - ms);
    }
    for (; j < j1; ++j) {
        const float x = col[j * ldl];
        if (x > m) {
            s *= (m == -INFINITY) ? 0.f : expf(m - x);
            m = x;
        }
        s += expf(x - ((m == -INFINITY) ? 0.f : m));
    }
    partials[(b * 2 + 0) * C + c] = m;
    partials[(b * 2 + 1) * C + c] = s;
}

// 32 columns x 32 block-slices per CTA: slice s reduces blocks s, s+32, ... (many independent loads in flight instead of
// one long dependent chain: 128 blocks at c4 are 4 per thread), then thread (c, 0) folds the slice results in slice
// order -- a fixed order, so the result is still independent of how the neurons were sharded.
constexpr int kLseSlices = 32;

__global__ void __launch_bounds__(32 * kLseSlices)
lse_combine_kernel(const float *__restrict__ partials, int64_t n_blocks, int C, double log_count,
                   float *__restrict__ prob_d, const int32_t *__restrict__ seg_tab, const double *__restrict__ seg_log) {
    pdl_enter();
    __shared__ float s_max[kLseSlices][33];
    __shared__ double s_sum[kLseSlices][33];
    const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    if (seg_tab) {       // segment (layer) blockIdx.y: its own run of blocks, its own neuron count, its own output row
        partials += int64_t(seg_tab[blockIdx.y * 2 + 0]) * 2 * C;
        n_blocks = seg_tab[blockIdx.y * 2 + 1];
        log_count = seg_log[blockIdx.y];
        prob_d += int64_t(blockIdx.y) * C;
    }
    float big = -INFINITY;
    if (c < C) {
#pragma unroll 4
        for (int64_t b = sl; b < n_blocks; b += kLseSlices) big = fmaxf(big, __ldg(partials + (b * 2) * C + c));
    }
    s_max[sl][cx] = big;
    __syncthreads();
    float all = s_max[0][cx];
#pragma unroll
    for (int q = 1; q < kLseSlices; ++q) all = fmaxf(all, s_max[q][cx]);
    const float bs = isinf(all) ? 0.f : all;
    // block weights exp(m_b - M) in fp32 (an fp64 exp dominated this kernel), accumulation in fp64
    double total = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int64_t b = sl; b < n_blocks; b += kLseSlices)
            total += double(__ldg(partials + (b * 2 + 1) * C + c)) * double(expf(__ldg(partials + (b * 2) * C + c) - bs));
    }
    s_sum[sl][cx] = total;
    __syncthreads();
    if (sl == 0 && c < C) {
        double t = s_sum[0][cx];
#pragma unroll
        for (int q = 1; q < kLseSlices; ++q) t += s_sum[q][cx];
        prob_d[c] = static_cast<float>(double(bs) + log(t) - log_count);
    }
}

// thread per concept, rows strided over blockIdx.y: no integer division, coalesced along the concept axis
__global__ void __launch_bounds__(256)
pmi_finalize_kernel(const float *__restrict__ L, int64_t ldl, int64_t K, int C, const float *__restrict__ prob_d,
                    float lam, float *__restrict__ out, int64_t ldo) {
    pdl_enter();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const float shift = __fmul_rn(lam, prob_d[c]);
    for (int64_t j = blockIdx.y; j < K; j += gridDim.y) out[j * ldo + c] = __fsub_rn(L[j * ldl + c], shift);
}

// several layers in one matrix: CTA (x, b) finalizes the rows of block b with its segment's log p(d)
__global__ void __launch_bounds__(256)
pmi_finalize_seg_kernel(const float *__restrict__ L, int64_t ldl, int C, const float *__restrict__ prob_d /*[n_seg][C]*/,
                        const int32_t *__restrict__ block_tab, float lam, float *__restrict__ out, int64_t ldo) {
    pdl_enter();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const int64_t j0 = block_tab[blockIdx.y * 3 + 0], j1 = j0 + block_tab[blockIdx.y * 3 + 1];
    const float shift = __fmul_rn(lam, prob_d[int64_t(block_tab[blockIdx.y * 3 + 2]) * C + c]);
    for (int64_t j = j0; j < j1; ++j) out[j * ldo + c] = __fsub_rn(L[j * ldl + c], shift);
}

// K3b fused with the score all-gather: the finalized slice is stored straight into the [K_total, C] score matrix
// of every GPU of the node (peer memory over NVLink / NVSwitch; the rank's own copy is one of the destinations).
// The slice is treated as a flat array so that every warp store is one aligned 512-byte run per destination, whatever
// C is; the concept index of a thread advances by a constant modulo C (no division in the loop).
constexpr int kMaxPeers = MCD_MAX_PEERS;
struct PeerDests {
    float *ptr[kMaxPeers];      // already offset to the first row of this rank's slice
};

__global__ void __launch_bounds__(256)
pmi_finalize_bcast_kernel(const float *__restrict__ L, int64_t total, int C, const float *__restrict__ prob_d, float lam,
                          PeerDests dst, int n_dst) {
    pdl_enter();
    const int64_t nvec = total / 4;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    int c = static_cast<int>((4 * e) % C);
    const int step_c = static_cast<int>((4 * stride) % C);
    for (; e < nvec; e += stride) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(L) + e);
        int c1 = c + 1 == C ? 0 : c + 1;
        int c2 = c1 + 1 == C ? 0 : c1 + 1;
        int c3 = c2 + 1 == C ? 0 : c2 + 1;
        v.x = __fsub_rn(v.x, __fmul_rn(lam, __ldg(prob_d + c)));
        v.y = __fsub_rn(v.y, __fmul_rn(lam, __ldg(prob_d + c1)));
        v.z = __fsub_rn(v.z, __fmul_rn(lam, __ldg(prob_d + c2)));
        v.w = __fsub_rn(v.w, __fmul_rn(lam, __ldg(prob_d + c3)));
        // statically indexed (a dynamic index would copy the pointer table to local memory)
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < n_dst) reinterpret_cast<float4 *>(dst.ptr[p])[e] = v;
        c += step_c;
        if (c >= C) c -= C;
    }
    // the last total % 4 elements
    if (blockIdx.x == 0 && threadIdx.x < total - nvec * 4) {
        const int64_t i = nvec * 4 + threadIdx.x;
        const float v = __fsub_rn(L[i], __fmul_rn(lam, prob_d[i % C]));
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < n_dst) dst.ptr[p][i] = v;
    }
}

// n floats to the same offset of several buffers (this rank's LSE partials into every peer's [blocks, 2, C] table through
// peer-mapped memory: replaces an NCCL all_gather of a few kilobytes, whose launch and protocol latency -- ~60 us -- is a
// tenth of a rank's whole step at 8 GPUs, by one small kernel + the symmetric-memory barrier).
__global__ void __launch_bounds__(256)
bcast_kernel(const float *__restrict__ src, int64_t n, PeerDests dst, int n_dst) {
    pdl_enter();
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(src + i);
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < n_dst) dst.ptr[p][i] = v;
    }
}

// K3b's finalize fused with the callers' per-neuron top concepts (describe_broad_neurons.py:101: torch.topk(sim, 10, 1);
// describe_clip_neurons.py:64: torch.max(sim, 1)): one warp per neuron row computes out = L - lam * log p(d) (the same
// two roundings as pmi_finalize_kernel), stores the row, and -- the values still in registers -- emits the row's t
// best (value, concept) pairs under the stated order (value desc, concept asc, NaN largest), so the [K, C] matrix is not
// read a second time.  row_seg (optional): the layer of every row when several layers share the matrix.
constexpr int kFinTopWarps = 8;

template <int PER>     // rows up to 32 * PER concepts
__global__ void __launch_bounds__(kFinTopWarps * 32)
pmi_finalize_topk_kernel(const float *__restrict__ L, int64_t ldl, int64_t K, int C, const float *__restrict__ prob_d,
                         const int32_t *__restrict__ row_seg, float lam, float *__restrict__ out, int64_t ldo, int t,
                         float *__restrict__ top_vals, int64_t *__restrict__ top_idx) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * kFinTopWarps + (threadIdx.x >> 5);
    if (row >= K) return;
    const float *src = L + row * ldl;
    const float *pd = prob_d + (row_seg ? int64_t(row_seg[row]) * C : 0);
    float *dst = out + row * ldo;
    unsigned long long w[PER];          // (ordered key, ~concept); 0 = no entry
    float v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        w[i] = 0ull;
        v[i] = 0.f;
        if (c < C) {
            v[i] = __fsub_rn(src[c], __fmul_rn(lam, pd[c]));
            dst[c] = v[i];
            w[i] = pack_key(ordered_key(v[i]), ~static_cast<uint32_t>(c));
        }
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
        // the winner's lane writes it (its value is in a register) and retires the entry: entries are distinct
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (w[i] == best) {
                top_idx[row * t + r] = static_cast<int64_t>(~static_cast<uint32_t>(best));
                top_vals[row * t + r] = v[i];
                w[i] = 0ull;
            }
    }
}

static int launch_finalize_topk(const float *L, int64_t ldl, int64_t K, int64_t C, const float *prob_d, const int32_t *row_seg,
                                float lam, float *out, int64_t ldo, int64_t t, float *top_vals, int64_t *top_idx,
                                cudaStream_t st) {
    if (C > 32 * 32 || t > 64 || t > C || t < 1) return MCD_ERR_UNSUPPORTED;
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(K, kFinTopWarps));
    const int Ci = static_cast<int>(C), ti = static_cast<int>(t);
    if (C <= 32 * 8)
        launch_pdl((pmi_finalize_topk_kernel<8>), dim3(grid), dim3(kFinTopWarps * 32), 0, st, L, ldl, K, Ci, prob_d, row_seg, lam, out, ldo, ti, top_vals, top_idx);
    else if (C <= 32 * 24)
        launch_pdl((pmi_finalize_topk_kernel<24>), dim3(grid), dim3(kFinTopWarps * 32), 0, st, L, ldl, K, Ci, prob_d, row_seg, lam, out, ldo, ti, top_vals, top_idx);
    else
        launch_pdl((pmi_finalize_topk_kernel<32>), dim3(grid), dim3(kFinTopWarps * 32), 0, st, L, ldl, K, Ci, prob_d, row_seg, lam, out, ldo, ti, top_vals, top_idx);
    return check_launch();
}

}  // namespace mcd

extern "C" int mcd_col_lse_partials_f32(const float *L, int64_t ldl, int64_t K, int64_t C, float *partials,
                                        mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials || K < 1 || C < 1 || ldl < C || C > (1 << 24)) return MCD_ERR_INVALID_ARGUMENT;
    const int64_t nb = ceil_div<int64_t>(K, MCD_LSE_BLOCK);
    if (nb > 65535) return MCD_ERR_UNSUPPORTED;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, kLseThreads)), static_cast<unsigned>(nb));
    launch_pdl((col_lse_partials_kernel), dim3(grid), dim3(kLseThreads), 0, static_cast<cudaStream_t>(stream), L, ldl, K, int(C), partials, nullptr);
    return check_launch();
}

extern "C" int mcd_pmi_finalize_f32(const float *L, int64_t ldl, int64_t K, int64_t C, const float *partials_all,
                                    int64_t n_blocks_total, int64_t K_total, float lam, float *prob_d_out,
                                    float *out, int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials_all || !prob_d_out || !out || K < 1 || C < 1 || ldl < C || ldo < C || n_blocks_total < 1 ||
        K_total < 1 || C > (1 << 24))
        return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    launch_pdl((lse_combine_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(C, 32))), dim3(32 * kLseSlices), 0, st, partials_all, n_blocks_total, int(C), log(double(K_total)), prob_d_out, nullptr, nullptr);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    int64_t rows = int64_t(num_sms()) * 16 / ceil_div<int64_t>(C, 256);
    if (rows > K) rows = K;
    if (rows > 65535) rows = 65535;
    if (rows < 1) rows = 1;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, 256)), static_cast<unsigned>(rows));
    launch_pdl((pmi_finalize_kernel), dim3(grid), dim3(256), 0, st, L, ldl, K, int(C), prob_d_out, lam, out, ldo);
    return check_launch();
}

extern "C" int mcd_pmi_finalize_topk_f32(const float *L, int64_t ldl, int64_t K, int64_t C, const float *partials_all,
                                         int64_t n_blocks_total, int64_t K_total, float lam, float *prob_d_out, float *out,
                                         int64_t ldo, int64_t t, float *top_vals_out, int64_t *top_idx_out,
                                         mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials_all || !prob_d_out || !out || !top_vals_out || !top_idx_out || K < 1 || C < 1 || ldl < C || ldo < C ||
        n_blocks_total < 1 || K_total < 1)
        return MCD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    launch_pdl((lse_combine_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(C, 32))), dim3(32 * kLseSlices), 0, st, partials_all, n_blocks_total, int(C), log(double(K_total)), prob_d_out, nullptr, nullptr);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    return launch_finalize_topk(L, ldl, K, C, prob_d_out, nullptr, lam, out, ldo, t, top_vals_out, top_idx_out, st);
}

extern "C" int mcd_pmi_finalize_bcast_f32(const float *L, int64_t K, int64_t C, const float *partials_all,
                                          int64_t n_blocks_total, int64_t K_total, float lam, float *prob_d_out,
                                          float *const *dest_bases, int n_dest, int64_t row_offset,
                                          mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials_all || !prob_d_out || !dest_bases || K < 1 || C < 1 || n_blocks_total < 1 || K_total < 1 ||
        C > (1 << 24) || n_dest < 1 || row_offset < 0 || row_offset + K > K_total)
        return MCD_ERR_INVALID_ARGUMENT;
    if (n_dest > kMaxPeers) return MCD_ERR_UNSUPPORTED;
    PeerDests dst;
    for (int p = 0; p < kMaxPeers; ++p) dst.ptr[p] = nullptr;
    for (int p = 0; p < n_dest; ++p) {
        if (!dest_bases[p]) return MCD_ERR_INVALID_ARGUMENT;
        dst.ptr[p] = dest_bases[p] + row_offset * C;
        if (reinterpret_cast<uintptr_t>(dst.ptr[p]) % 16 != 0) return MCD_ERR_UNSUPPORTED;   // 16-byte stores
    }
    if (reinterpret_cast<uintptr_t>(L) % 16 != 0) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    launch_pdl((lse_combine_kernel), dim3(static_cast<unsigned>(ceil_div<int64_t>(C, 32))), dim3(32 * kLseSlices), 0, st, partials_all, n_blocks_total, int(C), log(double(K_total)), prob_d_out, nullptr, nullptr);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    const int64_t total = K * C;
    int64_t blocks = ceil_div<int64_t>(total / 4 + 1, 256);
    const int64_t cap = int64_t(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    pmi_finalize_bcast_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(L, total, int(C), prob_d_out, lam, dst,
                                                                             n_dest);
    return check_launch();
}

// ---- several layers (segments of the neuron axis) in one matrix: SURVEY.md 8 f1 -------------------------------------
extern "C" int mcd_col_lse_partials_seg_f32(const float *L, int64_t ldl, int64_t C, const int32_t *block_tab,
                                            int64_t n_blocks, float *partials, mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials || !block_tab || C < 1 || ldl < C || C > (1 << 24) || n_blocks < 1) return MCD_ERR_INVALID_ARGUMENT;
    if (n_blocks > 65535) return MCD_ERR_UNSUPPORTED;
    dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(C, kLseThreads)), static_cast<unsigned>(n_blocks));
    launch_pdl((col_lse_partials_kernel), dim3(grid), dim3(kLseThreads), 0, static_cast<cudaStream_t>(stream), L, ldl, 0, int(C), partials, block_tab);
    return check_launch();
}

extern "C" int mcd_pmi_finalize_seg_f32(const float *L, int64_t ldl, int64_t C, const float *partials,
                                        const int32_t *block_tab, int64_t n_blocks, const int32_t *seg_tab,
                                        const double *seg_log_count, int64_t n_seg, float lam, float *prob_d_out,
                                        float *out, int64_t ldo, mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials || !block_tab || !seg_tab || !seg_log_count || !prob_d_out || !out || C < 1 || ldl < C || ldo < C ||
        C > (1 << 24) || n_blocks < 1 || n_seg < 1)
        return MCD_ERR_INVALID_ARGUMENT;
    if (n_blocks > 65535 || n_seg > 65535) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 cgrid(static_cast<unsigned>(ceil_div<int64_t>(C, 32)), static_cast<unsigned>(n_seg));
    launch_pdl((lse_combine_kernel), dim3(cgrid), dim3(32 * kLseSlices), 0, st, partials, 0, int(C), 0.0, prob_d_out, seg_tab, seg_log_count);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    dim3 fgrid(static_cast<unsigned>(ceil_div<int64_t>(C, 256)), static_cast<unsigned>(n_blocks));
    launch_pdl((pmi_finalize_seg_kernel), dim3(fgrid), dim3(256), 0, st, L, ldl, int(C), prob_d_out, block_tab, lam, out, ldo);
    return check_launch();
}

extern "C" int mcd_pmi_finalize_seg_topk_f32(const float *L, int64_t ldl, int64_t K, int64_t C, const float *partials,
                                             const int32_t *row_seg, int64_t n_blocks, const int32_t *seg_tab,
                                             const double *seg_log_count, int64_t n_seg, float lam, float *prob_d_out,
                                             float *out, int64_t ldo, int64_t t, float *top_vals_out,
                                             int64_t *top_idx_out, mcd_stream_t stream) {
    using namespace mcd;
    if (!L || !partials || !row_seg || !seg_tab || !seg_log_count || !prob_d_out || !out || !top_vals_out || !top_idx_out ||
        K < 1 || C < 1 || ldl < C || ldo < C || n_blocks < 1 || n_seg < 1)
        return MCD_ERR_INVALID_ARGUMENT;
    if (n_seg > 65535) return MCD_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 cgrid(static_cast<unsigned>(ceil_div<int64_t>(C, 32)), static_cast<unsigned>(n_seg));
    launch_pdl((lse_combine_kernel), dim3(cgrid), dim3(32 * kLseSlices), 0, st, partials, 0, int(C), 0.0, prob_d_out, seg_tab, seg_log_count);
    int rc = check_launch();
    if (rc != MCD_OK) return rc;
    return launch_finalize_topk(L, ldl, K, C, prob_d_out, row_seg, lam, out, ldo, t, top_vals_out, top_idx_out, st);
}

extern "C" int mcd_bcast_f32(const float *src, int64_t n, float *const *dest_bases, int n_dest, int64_t dest_offset,
                             mcd_stream_t stream) {
    using namespace mcd;
    if (!src || !dest_bases || n < 1 || n_dest < 1 || n_dest > kMaxPeers || dest_offset < 0) return MCD_ERR_INVALID_ARGUMENT;
    PeerDests dst;
    for (int p = 0; p < kMaxPeers; ++p) dst.ptr[p] = p < n_dest ? dest_bases[p] + dest_offset : nullptr;
    for (int p = 0; p < n_dest; ++p)
        if (!dest_bases[p]) return MCD_ERR_INVALID_ARGUMENT;
    int64_t blocks = ceil_div<int64_t>(n, 256);
    if (blocks > 4 * int64_t(num_sms())) blocks = 4 * int64_t(num_sms());
    launch_pdl((bcast_kernel), dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, n, dst, n_dest);
    return check_launch();
}
