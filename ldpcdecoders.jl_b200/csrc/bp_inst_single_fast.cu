// the node-parallel small-batch kernel (fast FP32 variant) in its own translation unit
#define BP_VARIANT 2
#include "bp_single.cuh"

namespace bp {

cudaError_t single_launch_2(int grid, int smem_bytes, cudaStream_t st, const SingleParams &p)
{
    auto k = bp_node_parallel_kernel<kSingleThreads>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, kSingleThreads, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace bp
