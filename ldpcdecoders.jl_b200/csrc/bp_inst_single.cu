// the node-parallel small-batch kernel (exact variant) in its own translation unit
#include "bp_single.cuh"

namespace bp {

cudaError_t single_launch(int grid, int smem_bytes, cudaStream_t st, const SingleParams &p)
{
    auto k = bp_node_parallel_kernel<kSingleThreads>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, kSingleThreads, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace bp
