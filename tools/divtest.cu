// Debug helper: lists operands where the fast node maps differ from the stock IEEE forms.
#include <cstdio>
#include "../ldpcdecoders.jl_b200/csrc/bp_math.cuh"
__global__ void k(unsigned long long n, unsigned long long* out, unsigned int* cnt)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t o[4];
        bp::philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 1u, 7u, 99u, 0u, o);
        const unsigned long long mant = (((unsigned long long)o[0] << 32) | o[1]) & 0x000fffffffffffffull;
        const unsigned kk = 1u + (o[2] % ((o[3] & 4u) ? 60u : 1000u));
        double mag = __longlong_as_double((long long)(((0x3ffull - kk) << 52) | mant));
        if (o[3] & 1u) mag = __dsub_rn(1.0, __longlong_as_double((long long)(((0x3ffull - (1u + o[2] % 53u)) << 52) | mant)));
        const double x = (o[3] & 2u) ? -mag : mag;
        if (!bp::rmap_envelope(x)) continue;
        const double f = bp::rmap_fast(x), g = bp::rmap(x);
        if (__double_as_longlong(f) != __double_as_longlong(g)) {
            unsigned int s = atomicAdd(cnt, 1u);
            if (s < 16) { out[3*s] = __double_as_longlong(x); out[3*s+1] = __double_as_longlong(f); out[3*s+2] = __double_as_longlong(g); }
        }
    }
}
int main(){
    unsigned long long* out; unsigned int* cnt;
    cudaMallocManaged(&out, 16*3*8); cudaMallocManaged(&cnt, 4); *cnt = 0;
    k<<<148*4,256>>>(1ull<<24, out, cnt);
    cudaDeviceSynchronize();
    printf("mismatches %u of %llu\n", *cnt, 1ull<<24);
    for (unsigned i = 0; i < (*cnt < 16 ? *cnt : 16); ++i) {
        double x, f, g; memcpy(&x,&out[3*i],8); memcpy(&f,&out[3*i+1],8); memcpy(&g,&out[3*i+2],8);
        printf("x=%.17g (%016llx) fast=%.17g (%016llx) ieee=%.17g (%016llx)\n", x, out[3*i], f, out[3*i+1], g, out[3*i+2]);
    }
    return 0;
}
