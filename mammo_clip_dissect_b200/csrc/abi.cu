// Library-level entry points of include/mcd_b200.h: version, errors, launch accounting, tunables.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace mcd {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int64_t> g_tunables[kNumTunables];
static std::atomic<int> g_num_sms{0};

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
int64_t tunable(int which) { return g_tunables[which].load(std::memory_order_relaxed); }

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
    static const char *names[] = {"topk_splits", "accum_tile", "topk_variant", "accum_unroll", "topk_cols", "topk_stages", "topk_occ", "gemm_variant", "topk_pre", "topk_small"};
    if (!name) return MCD_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 10; ++i)
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
    float *S = reinterpret_cast<float *>(w + l.s_off);
    int32_t *idx = reinterpret_cast<int32_t *>(w + l.idx_off);
    float *part = reinterpret_cast<float *>(w + l.part_off);
    float *prob_d = reinterpret_cast<float *>(w + l.probd_off);
    int rc = mcd_softmax_rows_f32(P, ldp, S, l.lds, N, C, a, stream);
    if (rc != MCD_OK) return rc;
    rc = mcd_topk_cols_f32(A, lda, N, K, k, nullptr, idx, nullptr, w + l.topk_off, workspace_bytes - l.topk_off, stream);
    if (rc != MCD_OK) return rc;
    rc = mcd_wpmi_accum_f32(S, l.lds, N, C, idx, K, k, p, min_prob, out, ldo, stream);
    if (rc != MCD_OK) return rc;
    rc = mcd_col_lse_partials_f32(out, ldo, K, C, part, stream);
    if (rc != MCD_OK) return rc;
    return mcd_pmi_finalize_f32(out, ldo, K, C, part, (K + MCD_LSE_BLOCK - 1) / MCD_LSE_BLOCK, K, lam, prob_d, out, ldo,
                                stream);
}
