// fp64_microbench.cu -- DFMA / MUFU.RCP64H latency and FP64 pipe throughput on the box's GPU (design input for the BP kernels).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64mb tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double *out, int n, long long *cyc, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void rcp_chain(double *out, int n, long long *cyc)
{
    double x = 1.5 + threadIdx.x * 1e-3;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        x = r;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
    long long h[1024];
    const int n = 4096;
    auto run = [&](auto kern, int chains, int blocks, int threads, const char *what) {
        kern<<<blocks, threads>>>(out, n, cyc, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(out, n, cyc, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, sizeof(long long) * (blocks < 1024 ? blocks : 1024), cudaMemcpyDeviceToHost);
        const double ops = double(n) * chains;
        printf("%-44s blocks %4d threads %4d: %.2f cycles per DFMA per warp (block 0), %.3f ms, %.2f T lane-ops/s\n", what, blocks, threads,
               h[0] / ops, ms, ops * blocks * threads / ms / 1e9);
    };
    run(dfma_chain<1>, 1, 1, 32, "dependent DFMA chain, 1 warp");
    run(dfma_chain<2>, 2, 1, 32, "2 independent chains, 1 warp");
    run(dfma_chain<4>, 4, 1, 32, "4 independent chains, 1 warp");
    run(dfma_chain<8>, 8, 1, 32, "8 independent chains, 1 warp");
    run(dfma_chain<1>, 1, 1, 128, "1 chain, 4 warps (one per SMSP)");
    run(dfma_chain<1>, 1, 1, 512, "1 chain, 16 warps (4 per SMSP)");
    run(dfma_chain<4>, 4, 1, 512, "4 chains, 16 warps (4 per SMSP)");
    run(dfma_chain<6>, 6, 1, 512, "6 chains, 16 warps (4 per SMSP)");
    run(dfma_chain<4>, 4, 148 * 2, 256, "4 chains, 2 x 256 threads per SM, full GPU");
    run(dfma_chain<8>, 8, 148 * 2, 256, "8 chains, 2 x 256 threads per SM, full GPU");
    run(dfma_chain<8>, 8, 148 * 2, 512, "8 chains, 2 x 512 threads per SM, full GPU");
    rcp_chain<<<1, 32>>>(out, n, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent MUFU.RCP64H chain, 1 warp: %.2f cycles per op\n", double(h[0]) / n);
    rcp_chain<<<1, 512>>>(out, n, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RCP64H, 16 warps: %.2f cycles per op per warp\n", double(h[0]) / n);
    return 0;
}
