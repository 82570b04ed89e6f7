// Internal (library-side) interface of the column top-k (K2) for callers that pipeline it by column chunks
// (abi.cu: mcd_pmi_logsums_f32 runs the gather / log-sum of chunk q under the scan of chunk q + 1).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mcd {

struct TopkPlan {
    int nstage, occ, splits, mpad;
    int64_t rows_per_split;
    size_t smem;
    size_t cand_bytes, kept_bytes;      // workspace: candidates [splits*k][K], then one kept region per scan warp
    // pre-threshold pass (single-split scans of long columns): every pre_stride-th row, k-th largest = pre_k
    int pre_stride, pre_k;
    int64_t pre_rows;
    size_t pre_bytes;                   // tau [K] floats, flags [ncb] ints, kept regions of the pre-pass
    // filter form (topk_filter.cuh): survivor lists instead of kept sets
    int filter;                         // 1: the call takes the filter path
    int f_cap, f_chunk_tiles, f_chunks, f_nstage, f_rows;
    size_t f_cnt_bytes, f_list_bytes;   // survivor counts [K] + item counters, lists [K][f_cap]
};

struct TopkFilterCall {
    CUtensorMap map_filter;             // [8 rows x 128 cols] boxes over A
    CUtensorMap map_scan;               // [64 rows x 32 cols] boxes (exact redo of flagged column groups)
    TopkPlan plan;
    const float *A;
    int64_t lda, N, K;
    int k;
    unsigned long long *cand;
    uint32_t *kept;
    float *tau;
    int *flags;
    uint32_t *tilemax;
    int *cnt;                           // [K] survivor counts, then kFMaxLaunches item counters
    unsigned long long *lists;
};

// Host only.  Returns MCD_OK and fills *c when the problem takes the filter path (long, TMA-aligned columns);
// MCD_ERR_UNSUPPORTED when it does not (the caller uses mcd_topk_cols_f32), other codes on bad arguments.
int topk_filter_prepare(const float *A, int64_t lda, int64_t N, int64_t K, int64_t k, void *workspace,
                        size_t workspace_bytes, TopkFilterCall *c);
// zero the counters, sample pass -> start threshold per column (all K columns)
int topk_filter_begin(const TopkFilterCall &c, cudaStream_t st);
// filter scan of columns [col0, col1) (col0 a multiple of 4, col1 a multiple of 256 or K); launch_id < 64 selects the
// launch's item counter
int topk_filter_scan(const TopkFilterCall &c, int64_t col0, int64_t col1, int launch_id, cudaStream_t st);
// select the top k of every column in [col0, col1) (col0 a multiple of 32), redo flagged groups exactly, emit [k, K] outputs
int topk_filter_finish(const TopkFilterCall &c, int64_t col0, int64_t col1, int64_t *idx64, int32_t *idx32, float *vals,
                       cudaStream_t st);

// K3 on a column range: idx points at the range's first column of the [k, idx_ld] index matrix
// probabilities: the caller vouches for S in [0, 1] (what the softmax kernels write): enables the grouped-log evaluation
int wpmi_accum_range(const float *S, int64_t lds, int64_t N, int64_t C, const int32_t *idx, int64_t idx_ld, int64_t K,
                     int64_t k, const float *p, float min_prob, float *L, int64_t ldl, cudaStream_t st, bool probabilities,
                     size_t pad_smem = 0);

}  // namespace mcd
