// Library-level entry points of include/mcd_b200.h: version, errors, launch accounting, tunables.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "topk_api.cuh"

namespace mcd {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int64_t> g_tunables[kNumTunables];
static std::atomic<int> g_num_sms{0};

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
int64_t tunable(int which) { return g_tunables[which].load(std::memory_order_relaxed); }

int debug_sync_check() {
    static const int on = [] {
        const char *e = getenv("MCD_DEBUG_SYNC");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    if (!on) return MCD_OK;
    const cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) return MCD_OK;
    fprintf(stderr, "[mcd] launch #%llu failed: %s\n", static_cast<unsigned long long>(g_launches.load()), cudaGetErrorString(err));
    return MCD_ERR_CUDA;
}

int num_sms() {
    int n = g_num_sms.load(std::memory_order_relaxed);
    if (n > 0) return n;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsB200;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsB200;
    g_num_sms.store(n, std::memory_order_relaxed);
    return n;
}

}  // namespace mcd

extern "C" {

int mcd_abi_version(void) { return MCD_ABI_VERSION; }

const char *mcd_strerror(int code) {
    switch (code) {
        case MCD_OK: return "ok";
        case MCD_ERR_INVALID_ARGUMENT: return "invalid argument";
        case MCD_ERR_UNSUPPORTED: return "unsupported size or layout";
        case MCD_ERR_WORKSPACE: return "workspace too small";
        case MCD_ERR_CUDA: return "CUDA launch failed";
        case MCD_ERR_NO_DEVICE: return "no sm_100 device";
        default: return "unknown error";
    }
}

#define MCD_STR2(x) #x
#define MCD_STR(x) MCD_STR2(x)
const char *mcd_build_info(void) {
    return "mcd_b200 abi " MCD_STR(MCD_ABI_VERSION) " sm_100a nvcc " MCD_STR(__CUDACC_VER_MAJOR__) "." MCD_STR(
        __CUDACC_VER_MINOR__) "." MCD_STR(__CUDACC_VER_BUILD__);
}

uint64_t mcd_launch_count(void) { return mcd::g_launches.load(std::memory_order_relaxed); }

int mcd_device_check(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MCD_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return MCD_ERR_NO_DEVICE;
    return major == 10 ? MCD_OK : MCD_ERR_NO_DEVICE;
}

int mcd_set_tunable(const char *name, int64_t value) {
    static const char *names[] = {"topk_splits", "accum_tile", "topk_variant", "accum_unroll", "topk_cols", "topk_stages", "topk_occ", "gemm_variant", "topk_pre", "topk_small", "topk_filter", "filter_stages", "filter_chunk_tiles", "pipe_chunks", "filter_order", "accum_pad_kb", "gemm_tiles_per_cta", "gemm_debug_terms"};
    constexpr int n_names = sizeof(names) / sizeof(names[0]);
    if (!name) return MCD_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_names; ++i)
        if (std::strcmp(name, names[i]) == 0) {
            mcd::g_tunables[i].store(value, std::memory_order_relaxed);
            return MCD_OK;
        }
    return MCD_ERR_INVALID_ARGUMENT;
}

}  // extern "C"

// ---- the whole soft_wpmi / wpmi call behind one entry point -----------------------------------------------------------
// K1b -> K2 -> K3 -> K3b with the intermediates in one caller-provided workspace: what similarity.soft_wpmi runs, as a
// single FFI call (a 500-neuron layer is launch-bound from Python: ~10 calls and ~6 allocations per layer otherwise).
namespace {
struct PmiLayout {
    size_t s_off, idx_off, part_off, probd_off, topk_off, total;
    int64_t lds;
};
bool pmi_layout(int64_t N, int64_t K, int64_t C, int64_t k, PmiLayout *l) {
    const size_t topk = mcd_topk_cols_workspace_bytes(N, K, k);
    if (topk == 0) return false;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    l->lds = (C + 31) / 32 * 32;                       // rows of S start on 128-byte boundaries
    l->s_off = 0;
    l->idx_off = up(size_t(N) * size_t(l->lds) * 4);
    l->part_off = l->idx_off + up(size_t(k) * size_t(K) * 4);
    l->probd_off = l->part_off + up(size_t((K + MCD_LSE_BLOCK - 1) / MCD_LSE_BLOCK) * 2 * size_t(C) * 4);
    l->topk_off = l->probd_off + up(size_t(C) * 4);
    l->total = l->topk_off + up(topk);
    return true;
}
}  // namespace

extern "C" size_t mcd_pmi_scores_workspace_bytes(int64_t N, int64_t K, int64_t C, int64_t k) {
    PmiLayout l;
    if (N < 1 || K < 1 || C < 1 || k < 1 || k > N || !pmi_layout(N, K, C, k, &l)) return 0;
    return l.total;
}

namespace {

// ---- column-chunk pipeline -------------------------------------------------------------------------------------------
// The filter scan of K2 is HBM-bound and leaves most issue slots of an SM idle; the gather / log-sum of K3 is L2-bound.
// With the neurons cut into a few column chunks, chunk q's select + K3 + LSE partials run on a side stream under the scan
// of chunk q + 1 (and the softmax under the sample pass and the first scan), so that only the last, smallest chunk's K3
// is exposed.  The side stream and the events are created once per device, on first use.
constexpr int kMaxPipeChunks = 8;
struct PipeRes {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, scanned[kMaxPipeChunks] = {};
    bool ok = false, tried = false;
};
std::mutex g_pipe_mu;
PipeRes g_pipe[64];

PipeRes *pipe_resources() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pipe_mu);
    PipeRes &r = g_pipe[dev];
    if (!r.tried) {
        r.tried = true;
        bool ok = cudaStreamCreateWithFlags(&r.side, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&r.join, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < kMaxPipeChunks; ++i)
            ok = cudaEventCreateWithFlags(&r.scanned[i], cudaEventDisableTiming) == cudaSuccess;
        r.ok = ok;
    }
    return r.ok ? &r : nullptr;
}

// chunk boundaries (multiples of the 256-neuron LSE block) with geometrically shrinking widths: chunk q's K3 has to fit
// under the scan of chunk q + 1, and the last chunk's K3 is the exposed tail
int pipe_bounds(int64_t K, int64_t *bounds) {
    int64_t q = mcd::tunable(mcd::kPipeChunks);
    const int64_t units = (K + MCD_LSE_BLOCK - 1) / MCD_LSE_BLOCK;
    if (q <= 0) q = 1;
    if (q > kMaxPipeChunks) q = kMaxPipeChunks;
    if (q > units) q = units;
    bounds[0] = 0;
    if (q <= 1) {
        bounds[1] = K;
        return 1;
    }
    double w[kMaxPipeChunks], sum = 0.0, x = 1.0;
    for (int i = 0; i < q; ++i, x *= 0.72) sum += (w[i] = x);
    double acc = 0.0;
    int n = 0;
    for (int i = 0; i < q; ++i) {
        acc += w[i] / sum;
        int64_t b = i == q - 1 ? units : static_cast<int64_t>(acc * double(units) + 0.5);
        if (b > units) b = units;
        if (b * MCD_LSE_BLOCK > bounds[n]) bounds[++n] = b * MCD_LSE_BLOCK < K ? b * MCD_LSE_BLOCK : K;
    }
    bounds[n] = K;
    return n;
}

// softmax -> column top-k -> gather / log-sum -> 256-neuron block partials; L [K, C] and partials are outputs
int pmi_logsums(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K, int64_t C, int64_t k,
                float a, const float *p, float min_prob, float *L, int64_t ldl, float *part, const PmiLayout &l, char *w,
                size_t workspace_bytes, mcd_stream_t stream) {
    using namespace mcd;
    float *S = reinterpret_cast<float *>(w + l.s_off);
    int32_t *idx = reinterpret_cast<int32_t *>(w + l.idx_off);
    cudaStream_t main = static_cast<cudaStream_t>(stream);
    TopkFilterCall call;
    int64_t bounds[kMaxPipeChunks + 1];
    const int nq = pipe_bounds(K, bounds);
    // tunable pipe_chunks: 0 (default) / -1 = one stream, stage after stage; n > 0 = n column chunks with a side stream
    // (1: only the softmax runs beside the sample pass and the scan).  Measured at c4 on B200 (profiles/r2_k2_history.md):
    // every overlapped variant LOSES to the plain sequence (3.19 ms; 2 chunks 3.46, 4 chunks 3.62) -- K3's CTAs take
    // resident-warp slots and L2 bandwidth from the scan, which needs >= 9 resident warps per SM to saturate HBM, and
    // each chunk re-reads the concept-tile slices of S from DRAM -- so the pipeline stays an option, not the default.
    // What IS worth a second stream: the softmax (K1b) beside the sample pass and the scan when the neuron shard is narrow
    // (a rank of the 8-GPU call: 4096 neurons) -- the softmax over all N images does not shrink with the shard and would
    // otherwise be a third of the rank's step; at full width both are HBM-bound and the overlap is neutral.
    // (At full width the softmax was also tried beside the list SELECT -- instruction-bound, DRAM at 16 % -- and split
    // between the threshold select and the list select: 2.98 / 3.00 ms against 2.98 ms for the plain sequence on the same
    // box, i.e. nothing: both kernels are issue-heavy.  Removed.)
    const bool side_softmax = tunable(kPipeChunks) == 0 && N * C >= (int64_t(1) << 23) && K <= 16384;
    PipeRes *pr = (tunable(kPipeChunks) > 0 || side_softmax) ? pipe_resources() : nullptr;
    int rc = pr ? topk_filter_prepare(A, lda, N, K, k, w + l.topk_off, workspace_bytes - l.topk_off, &call) : MCD_ERR_UNSUPPORTED;
    if (rc != MCD_OK && rc != MCD_ERR_UNSUPPORTED) return rc;
    if (rc == MCD_ERR_UNSUPPORTED) {
        // one stream, stage after stage
        rc = mcd_softmax_rows_f32(P, ldp, S, l.lds, N, C, a, stream);
        if (rc != MCD_OK) return rc;
        rc = mcd_topk_cols_f32(A, lda, N, K, k, nullptr, idx, nullptr, w + l.topk_off, workspace_bytes - l.topk_off, stream);
        if (rc != MCD_OK) return rc;
        rc = mcd_wpmi_accum_prob_f32(S, l.lds, N, C, idx, K, k, p, min_prob, L, ldl, stream);      // S is our own softmax
        if (rc != MCD_OK) return rc;
        return mcd_col_lse_partials_f32(L, ldl, K, C, part, stream);
    }
    cudaStream_t side = pr->side;
    if (cudaEventRecord(pr->fork, main) != cudaSuccess || cudaStreamWaitEvent(side, pr->fork, 0) != cudaSuccess) return MCD_ERR_CUDA;
    rc = mcd_softmax_rows_f32(P, ldp, S, l.lds, N, C, a, side);
    if (rc != MCD_OK) return rc;
    rc = topk_filter_begin(call, main);
    if (rc != MCD_OK) return rc;
    for (int q = 0; q < nq; ++q) {
        const int64_t c0 = bounds[q], c1 = bounds[q + 1];
        rc = topk_filter_scan(call, c0, c1, q, main);
        if (rc != MCD_OK) return rc;
        if (cudaEventRecord(pr->scanned[q], main) != cudaSuccess || cudaStreamWaitEvent(side, pr->scanned[q], 0) != cudaSuccess)
            return MCD_ERR_CUDA;
        rc = topk_filter_finish(call, c0, c1, nullptr, idx, nullptr, side);
        if (rc != MCD_OK) return rc;
        // all chunks but the last run beside the scan of the next chunk: pad their shared memory (tunable accum_pad_kb) so
        // that only one or two of K3's CTAs fit next to the scan's resident warps on an SM -- the scan needs >= 9 of them
        const size_t pad = q + 1 < nq ? size_t(tunable(kAccumPadKb)) << 10 : 0;
        rc = wpmi_accum_range(S, l.lds, N, C, idx + c0, K, c1 - c0, k, p, min_prob, L + c0 * ldl, ldl, side, true, pad);
        if (rc != MCD_OK) return rc;
        rc = mcd_col_lse_partials_f32(L + c0 * ldl, ldl, c1 - c0, C, part + (c0 / MCD_LSE_BLOCK) * 2 * C, side);
        if (rc != MCD_OK) return rc;
    }
    if (cudaEventRecord(pr->join, side) != cudaSuccess || cudaStreamWaitEvent(main, pr->join, 0) != cudaSuccess) return MCD_ERR_CUDA;
    return MCD_OK;
}

}  // namespace

extern "C" int mcd_pmi_logsums_f32(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K,
                                   int64_t C, int64_t k, float a, const float *p, float min_prob, float *L, int64_t ldl,
                                   float *partials, void *workspace, size_t workspace_bytes, mcd_stream_t stream) {
    if (!P || !A || !L || !partials || !workspace || N < 1 || K < 1 || C < 1 || k < 1 || k > N || ldp < C || lda < K || ldl < C)
        return MCD_ERR_INVALID_ARGUMENT;
    PmiLayout l;
    if (!pmi_layout(N, K, C, k, &l)) return MCD_ERR_UNSUPPORTED;
    if (workspace_bytes < l.total) return MCD_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return MCD_ERR_INVALID_ARGUMENT;
    return pmi_logsums(P, ldp, A, lda, N, K, C, k, a, p, min_prob, L, ldl, partials, l, static_cast<char *>(workspace),
                       workspace_bytes, stream);
}

extern "C" int mcd_pmi_scores_f32(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K,
                                  int64_t C, int64_t k, float a, float lam, const float *p, float min_prob, float *out,
                                  int64_t ldo, void *workspace, size_t workspace_bytes, mcd_stream_t stream) {
    if (!P || !A || !out || !workspace || N < 1 || K < 1 || C < 1 || k < 1 || k > N || ldp < C || lda < K || ldo < C)
        return MCD_ERR_INVALID_ARGUMENT;
    PmiLayout l;
    if (!pmi_layout(N, K, C, k, &l)) return MCD_ERR_UNSUPPORTED;
    if (workspace_bytes < l.total) return MCD_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return MCD_ERR_INVALID_ARGUMENT;
    char *w = static_cast<char *>(workspace);
    float *part = reinterpret_cast<float *>(w + l.part_off);
    float *prob_d = reinterpret_cast<float *>(w + l.probd_off);
    int rc = pmi_logsums(P, ldp, A, lda, N, K, C, k, a, p, min_prob, out, ldo, part, l, w, workspace_bytes, stream);
    if (rc != MCD_OK) return rc;
    return mcd_pmi_finalize_f32(out, ldo, K, C, part, (K + MCD_LSE_BLOCK - 1) / MCD_LSE_BLOCK, K, lam, prob_d, out, ldo,
                                stream);
}
