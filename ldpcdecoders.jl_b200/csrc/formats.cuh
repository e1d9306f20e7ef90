// formats.cuh -- boundary formats <-> native packed rows, synthetic sampler, scoring.
//
// Native device format (LDPCB200_FMT_PACKED32): one row of ceil(rows/32) uint32 per syndrome,
// bit r%32 of word r/32.  The kernels here convert the host matrices batchdecode! is called
// with (/root/reference/src/decoders/belief_propagation.jl:220-231: Matrix{Int}, BitMatrix,
// Matrix{Bool}; outputs via `errors[:, i] .= guess`) on the device, so PCIe only carries the
// caller's own bytes.
#pragma once
#include "bp_math.cuh"

namespace bp {

// ---- column-major element matrices -> packed rows -------------------------------------------
template <typename T>
__global__ void pack_elems(const T *__restrict__ src, long long ld, int rows, int RW, long long B,
                           uint32_t *__restrict__ dst)
{
    const long long total = B * RW;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = t / RW;
        const int w = static_cast<int>(t - b * RW);
        const T *col = src + b * ld + w * 32;
        const int nb = min(32, rows - w * 32);
        uint32_t v = 0;
        for (int k = 0; k < nb; ++k) v |= (static_cast<uint32_t>(col[k] != T(0)) & 1u) << k;
        dst[t] = v;
    }
}

template <typename T>
__global__ void unpack_elems(const uint32_t *__restrict__ src, int rows, int RW, long long B,
                             T *__restrict__ dst, long long ld)
{
    const long long total = B * RW;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = t / RW;
        const int w = static_cast<int>(t - b * RW);
        T *col = dst + b * ld + w * 32;
        const int nb = min(32, rows - w * 32);
        const uint32_t v = src[t];
        for (int k = 0; k < nb; ++k) col[k] = static_cast<T>((v >> k) & 1u);
    }
}

// ---- Julia BitMatrix bit stream (bit c*rows + r) <-> packed rows -------------------------------
// `first_bit` = stream bit index of this chunk's first column inside `stream` (multiple of 32).
__device__ __forceinline__ uint32_t stream_bits32(const uint32_t *stream, long long nwords, long long bit)
{
    const long long w = bit >> 5;
    const uint32_t lo = (w < nwords) ? stream[w] : 0u;
    const uint32_t hi = (w + 1 < nwords) ? stream[w + 1] : 0u;
    return __funnelshift_r(lo, hi, static_cast<uint32_t>(bit & 31));
}

__global__ void pack_bits(const uint32_t *__restrict__ stream, long long nwords, int rows, int RW, long long B,
                          uint32_t *__restrict__ dst)
{
    const long long total = B * RW;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = t / RW;
        const int w = static_cast<int>(t - b * RW);
        const int nb = min(32, rows - w * 32);
        uint32_t v = stream_bits32(stream, nwords, b * rows + static_cast<long long>(w) * 32);
        if (nb < 32) v &= (1u << nb) - 1u;
        dst[t] = v;
    }
}

// bits r .. r+take-1 (take <= 32) of a packed row
__device__ __forceinline__ uint32_t row_bits(const uint32_t *row, int RW, int r, int take)
{
    const int w = r >> 5;
    const uint32_t lo = row[w];
    const uint32_t hi = (w + 1 < RW) ? row[w + 1] : 0u;
    uint32_t v = __funnelshift_r(lo, hi, static_cast<uint32_t>(r & 31));
    if (take < 32) v &= (1u << take) - 1u;
    return v;
}

__global__ void unpack_bits(const uint32_t *__restrict__ src, int rows, int RW, long long B,
                            uint32_t *__restrict__ stream, long long nwords)
{
    const long long total_bits = B * rows;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < nwords;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long bit = t << 5;
        uint32_t v = 0;
        int filled = 0;
        while (filled < 32 && bit < total_bits) {
            const long long c = bit / rows;
            const int r = static_cast<int>(bit - c * rows);
            const int take = min(32 - filled, rows - r);
            v |= row_bits(src + c * RW, RW, r, take) << filled;
            filled += take;
            bit += take;
        }
        stream[t] = v;
    }
}

// ---- posterior ratio rows [B][n] are already the column-major n x B matrix: plain copy --------

// ---- synthetic inputs, same stream as oracle/bp_oracle.c:bp_oracle_sample -----
__global__ void sample_errors(int n, int NW, long long B, long long first, uint32_t k0, uint32_t k1,
                              uint32_t thr, uint32_t *__restrict__ err_words)
{
    const long long total = B * NW;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = t / NW;
        const int w = static_cast<int>(t - b * NW);
        const unsigned long long gb = static_cast<unsigned long long>(first + b);
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int g = w * 8 + q;                 // bits 4g .. 4g+3
            if (4 * g < n) {
                uint32_t o[4];
                philox4x32_10(static_cast<uint32_t>(gb), static_cast<uint32_t>(gb >> 32),
                              static_cast<uint32_t>(g), 0u, k0, k1, o);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * g + i < n && o[i] < thr) v |= 1u << (4 * q + i);
            }
        }
        err_words[t] = v;
    }
}

// syn_words must be zero on entry; scatters the checks of every set error bit.
__global__ void syndrome_of(const int *__restrict__ colptr, const int *__restrict__ ve_chk, int NW, int SW,
                            long long B, const uint32_t *__restrict__ err_words, uint32_t *__restrict__ syn_words)
{
    const long long total = B * NW;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = t / NW;
        const int w = static_cast<int>(t - b * NW);
        uint32_t v = err_words[t];
        while (v) {
            const int j = w * 32 + (__ffs(v) - 1);
            v &= v - 1;
            for (int e = colptr[j]; e < colptr[j + 1]; ++e) {
                const int chk = ve_chk[e];
                atomicXor(syn_words + b * SW + (chk >> 5), 1u << (chk & 31));
            }
        }
    }
}

// out[0] += rows with decoded == truth ; out[1] += rows with H*decoded == syndrome.
// One warp per row.
__global__ void score_rows(const int *__restrict__ rowptr_unused, const int *__restrict__ colptr,
                           const int *__restrict__ ve_chk, int NW, int SW, long long B,
                           const uint32_t *__restrict__ truth, const uint32_t *__restrict__ dec,
                           const uint32_t *__restrict__ syn, uint32_t *__restrict__ scratch_syn,
                           unsigned long long *out)
{
    (void)rowptr_unused;
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    unsigned long long exact = 0, consistent = 0;
    for (long long b = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        bool same = true;
        for (int w = lane; w < NW; w += 32) same &= truth[b * NW + w] == dec[b * NW + w];
        same = __all_sync(0xffffffffu, same);
        // scratch row = syndrome of the decoded error
        uint32_t *row = scratch_syn + b * SW;
        for (int w = lane; w < SW; w += 32) row[w] = 0u;
        __syncwarp();
        for (int w = lane; w < NW; w += 32) {
            uint32_t v = dec[b * NW + w];
            while (v) {
                const int j = w * 32 + (__ffs(v) - 1);
                v &= v - 1;
                for (int e = colptr[j]; e < colptr[j + 1]; ++e) {
                    const int chk = ve_chk[e];
                    atomicXor(row + (chk >> 5), 1u << (chk & 31));
                }
            }
        }
        __syncwarp();
        bool ok = true;
        for (int w = lane; w < SW; w += 32) ok &= row[w] == syn[b * SW + w];
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) { exact += same; consistent += ok; }
    }
    if (lane == 0 && (exact | consistent)) {
        atomicAdd(out + 0, exact);
        atomicAdd(out + 1, consistent);
    }
}

// Scoring with logical operators (pattern of test/test_bp_decoder.jl:19-30 with the comparison a QEC
// caller really wants): r = truth xor decoded is a failure when it is detectable (H*decoded != syndrome) or acts on the
// logical qubits (L*r != 0, lmask[j] = column j of the k x n matrix L, k <= 64, as a bit mask).  Without logical
// operators (lmask == nullptr) a failure is `decoded != truth`, the reference test's own criterion.
// out[0] += exact matches, out[1] += rows with H*decoded == syndrome, out[2] += failures, out[3] += residual weight.
// One warp per row.
__global__ void score_rows_logical(const int *__restrict__ colptr, const int *__restrict__ ve_chk, int NW, int SW, long long B,
                                   const uint32_t *__restrict__ truth, const uint32_t *__restrict__ dec,
                                   const uint32_t *__restrict__ syn, uint32_t *__restrict__ scratch_syn,
                                   const unsigned long long *__restrict__ lmask, unsigned long long *out)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    unsigned long long exact = 0, consistent = 0, failures = 0, weight = 0;
    for (long long b = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        uint32_t *row = scratch_syn + b * SW;
        for (int w = lane; w < SW; w += 32) row[w] = 0u;
        __syncwarp();
        bool same = true;
        unsigned long long lacc = 0;
        int wt = 0;
        for (int w = lane; w < NW; w += 32) {
            uint32_t v = dec[b * NW + w];
            uint32_t r = v ^ truth[b * NW + w];
            same &= r == 0u;
            wt += __popc(r);
            while (v) {                                   // syndrome of the decoded error
                const int j = w * 32 + (__ffs(v) - 1);
                v &= v - 1;
                for (int e = colptr[j]; e < colptr[j + 1]; ++e) {
                    const int chk = ve_chk[e];
                    atomicXor(row + (chk >> 5), 1u << (chk & 31));
                }
            }
            if (lmask)
                while (r) {                               // action of the residual on the logical qubits
                    lacc ^= lmask[w * 32 + (__ffs(r) - 1)];
                    r &= r - 1;
                }
        }
        same = __all_sync(0xffffffffu, same);
        for (int o = 16; o > 0; o >>= 1) {
            lacc ^= __shfl_xor_sync(0xffffffffu, lacc, o);
            wt += __shfl_xor_sync(0xffffffffu, wt, o);
        }
        __syncwarp();
        bool ok = true;
        for (int w = lane; w < SW; w += 32) ok &= row[w] == syn[b * SW + w];
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) {
            exact += same; consistent += ok; weight += wt;
            failures += lmask ? (!ok || lacc != 0ull) : !same;
        }
    }
    if (lane == 0) {
        if (exact) atomicAdd(out + 0, exact);
        if (consistent) atomicAdd(out + 1, consistent);
        if (failures) atomicAdd(out + 2, failures);
        if (weight) atomicAdd(out + 3, weight);
    }
}

}  // namespace bp
