// bp_single.h -- host-visible part of the node-parallel small-batch kernel (bp_single.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bp {

struct SingleParams {
    int s, n, E, SW, NW, max_iters, early_stop, regular_p0, ratio_last_only;
    double p0;         // prior ratio p/(1-p); min-sum variant: prior log-likelihood ratio
    double check_aux;  // min-sum variant: normalisation factor
    long long B;
    const int *rowptr, *colptr, *ve_slot, *ve_chk;
    const uint32_t *syn_words;
    uint32_t *err_words;
    uint8_t *conv;
    int32_t *iters;
    double *ratio;
    unsigned long long *counters;
    int off_syn, off_resid, off_dec;      // shared-memory byte offsets behind the E message slots
};

constexpr int kSingleThreads = 512;
cudaError_t single_launch_0(int grid, int smem_bytes, cudaStream_t st, const SingleParams &p);   // exact variant
cudaError_t single_launch_1(int grid, int smem_bytes, cudaStream_t st, const SingleParams &p);   // min-sum variant
cudaError_t single_launch_2(int grid, int smem_bytes, cudaStream_t st, const SingleParams &p);   // fast FP32 variant

}  // namespace bp
