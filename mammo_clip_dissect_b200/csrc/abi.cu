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
