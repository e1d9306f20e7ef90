// bp_single.cuh -- node-parallel BP kernel for SMALL batches (decode! of one syndrome, a handful of columns).
//
// The persistent kernel (bp_kernel.cuh) maps one syndrome to one lane: a batch of 1 keeps 1 lane of 1 SM busy
// for max_iters * (s + n) sequential node updates.  Here one CTA decodes one syndrome and its threads own
// the nodes: thread t updates checks t, t+T, ... then variables t, t+T, ... of
// decode!(::BeliefPropagationDecoder, syndrome)  (/root/reference/src/decoders/belief_propagation.jl:121-188).
// Same arithmetic (bp_math.cuh: check_update<D>, var_update<D>, decide), same storage (one in-place message
// array in check-major edge order, in shared memory), same incremental residual syndrome for the :180-184
// early stop, so results are bit-identical to the persistent kernel and to the oracle.  Graph tables are in
// the ORIGINAL node order (no degree sorting needed: nodes are not walked in lock step).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef BP_VARIANT
#define BP_VARIANT 0
#endif
#include "bp_math.cuh"
#include "bp_single.h"

namespace bp {
inline namespace BP_VNS {

template <int T>
__global__ void __launch_bounds__(T, 1) bp_node_parallel_kernel(const SingleParams p)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double *msg = reinterpret_cast<double *>(sm_raw);
    uint32_t *syn = reinterpret_cast<uint32_t *>(sm_raw + p.off_syn);
    uint32_t *resid = reinterpret_cast<uint32_t *>(sm_raw + p.off_resid);
    uint32_t *dec = reinterpret_cast<uint32_t *>(sm_raw + p.off_dec);      // decisions, bit-packed like the output row
    const int tid = threadIdx.x;
    const double p0 = p.p0;
    const bool regular_p0 = p.regular_p0 != 0;

    for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int w = tid; w < p.SW; w += T) { const uint32_t v = p.syn_words[b * p.SW + w]; syn[w] = v; resid[w] = v; }
        for (int w = tid; w < p.NW; w += T) dec[w] = 0u;                   // err .= 0 (reset!, :89)
        __syncthreads();
        bool conv = false;
        int iter = 0;
        while (iter < p.max_iters) {                                       // :134
            const bool fresh = iter == 0;                                  // messages still hold the prior (:127-131)
            // ---- check pass (:135-150)
            for (int i = tid; i < p.s; i += T) {
                const int rp = p.rowptr[i], deg = p.rowptr[i + 1] - rp;
                const bool neg = (syn[i >> 5] >> (i & 31)) & 1u;
#define BP_CASE(D)                                                                    \
    {                                                                                 \
        double m[D];                                                                  \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = fresh ? p0 : msg[rp + k]; \
        check_update<D>(m, neg, p.check_aux);                                         \
        _Pragma("unroll") for (int k = 0; k < D; ++k) msg[rp + k] = m[k];             \
    }
                BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
            }
            __syncthreads();
            // ---- variable pass, decision, residual syndrome (:152-184)
            const bool wr = p.ratio != nullptr && (!p.ratio_last_only || iter + 1 >= p.max_iters);
            for (int j = tid; j < p.n; j += T) {
                const int cp = p.colptr[j], deg = p.colptr[j + 1] - cp;
                double R = p0;
#define BP_CASE(D)                                                                    \
    {                                                                                 \
        int sl[D];                                                                    \
        double m[D];                                                                  \
        _Pragma("unroll") for (int k = 0; k < D; ++k) sl[k] = p.ve_slot[cp + k];      \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = msg[sl[k]];              \
        R = var_update<D>(m, p0, regular_p0);                                         \
        _Pragma("unroll") for (int k = 0; k < D; ++k) msg[sl[k]] = m[k];              \
    }
                BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
                if (wr) p.ratio[b * p.n + j] = R;
                const uint32_t bit = decide(R) ? 1u : 0u;                  // tie -> 1 (:164)
                if (bit != ((dec[j >> 5] >> (j & 31)) & 1u)) {
                    atomicXor(&dec[j >> 5], 1u << (j & 31));
                    for (int e = cp; e < cp + deg; ++e) {
                        const int c = p.ve_chk[e];
                        atomicXor(&resid[c >> 5], 1u << (c & 31));
                    }
                }
            }
            __syncthreads();
            bool nz = false;
            for (int w = tid; w < p.SW; w += T) nz |= resid[w] != 0u;
            conv = __syncthreads_or(nz) == 0;                              // H*err mod 2 == syndrome (:180-181)
            ++iter;
            if (p.early_stop && conv) break;                               // :182-184
        }
        for (int w = tid; w < p.NW; w += T) p.err_words[b * p.NW + w] = dec[w];
        if (tid == 0) {
            p.conv[b] = conv ? 1 : 0;
            if (p.iters) p.iters[b] = iter;
            if (p.counters) {
                atomicAdd(&p.counters[0], 1ull);
                atomicAdd(&p.counters[1], conv ? 1ull : 0ull);
                atomicAdd(&p.counters[2], static_cast<unsigned long long>(iter));
            }
        }
        __syncthreads();
    }
}

// ---- the same decode spread over the WHOLE GRID (cooperative launch): one syndrome at a time, threads of all CTAs over
// the nodes, messages in global memory (they stay in L2: E * 8 bytes), two grid-wide barriers per iteration.  For a lone
// decode! / a handful of columns on codes whose messages do not fit in one SM's shared memory (n = 100k: 2.4 MB), where
// the persistent kernel would walk all E edges on a single lane.  Same node updates, same order of operations inside a
// node, same incremental residual -- here with a global count of unsatisfied checks kept by the returning atomics of the
// flipped decisions -- so results are bit-identical to the other kernels and the oracle.
//   work: [b & 1] unsatisfied checks of syndrome b (two counters, so that the next syndrome's count can be built while the
//   last readers of this one are still around); msg [E] doubles; resid [SW]; dec [NW] words
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nblocks, unsigned int &epoch)
{
    // sense-free counting barrier on a monotonically increasing counter: all CTAs are co-resident (cooperative launch)
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned int target = epoch * nblocks;
        while (*reinterpret_cast<volatile unsigned int *>(bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <int T>
__global__ void __launch_bounds__(T, 1) bp_grid_kernel(const GridParams p)
{
    const long long tid = static_cast<long long>(blockIdx.x) * T + threadIdx.x;
    const long long NT = static_cast<long long>(gridDim.x) * T;
    const double p0 = p.p0;
    const bool regular_p0 = p.regular_p0 != 0;
    double *msg = p.msg;
    unsigned int epoch = 0;
    for (long long b = 0; b < p.B; ++b) {
        int *nnz = p.work + (b & 1);
        if (tid == 0) p.work[(b + 1) & 1] = 0;        // the previous syndrome's counter: everyone left it at the last barrier
        const uint32_t *srow = p.syn_words + b * p.SW;
        {   // residual = syndrome, decisions = 0 (reset!, :89), unsatisfied count = weight of the syndrome
            int cnt = 0;
            for (long long w = tid; w < p.SW; w += NT) { const uint32_t v = srow[w]; p.resid[w] = v; cnt += __popc(v); }
            for (long long w = tid; w < p.NW; w += NT) p.dec[w] = 0u;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nnz, cnt);
        }
        grid_barrier(p.bar, gridDim.x, epoch);
        bool conv = false;
        int iter = 0;
        while (iter < p.max_iters) {                                       // :134
            const bool fresh = iter == 0;
            // ---- check pass (:135-150)
            for (long long i = tid; i < p.s; i += NT) {
                const int rp = p.rowptr[i], deg = p.rowptr[i + 1] - rp;
                const bool neg = (srow[i >> 5] >> (i & 31)) & 1u;
#define BP_CASE(D)                                                                            \
    {                                                                                         \
        double m[D];                                                                          \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = fresh ? p0 : __ldcg(msg + rp + k); \
        check_update<D>(m, neg, p.check_aux);                                                 \
        _Pragma("unroll") for (int k = 0; k < D; ++k) __stcg(msg + rp + k, m[k]);             \
    }
                BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
            }
            grid_barrier(p.bar, gridDim.x, epoch);
            // ---- variable pass, decision, residual syndrome (:152-184)
            const bool wr = p.ratio != nullptr && (!p.ratio_last_only || iter + 1 >= p.max_iters);
            int dn = 0;                                                    // change of the unsatisfied count by this thread's flips
            for (long long j = tid; j < p.n; j += NT) {
                const int cp = p.colptr[j], deg = p.colptr[j + 1] - cp;
                double R = p0;
#define BP_CASE(D)                                                                            \
    {                                                                                         \
        int sl[D];                                                                            \
        double m[D];                                                                          \
        _Pragma("unroll") for (int k = 0; k < D; ++k) sl[k] = p.ve_slot[cp + k];              \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = __ldcg(msg + sl[k]);             \
        R = var_update<D>(m, p0, regular_p0);                                                 \
        _Pragma("unroll") for (int k = 0; k < D; ++k) __stcg(msg + sl[k], m[k]);              \
    }
                BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
                if (wr) p.ratio[b * p.n + j] = R;
                const uint32_t bit = decide(R) ? 1u : 0u;                  // tie -> 1 (:164)
                // (bit j of dec is only ever touched by this thread, but the word is shared: atomics)
                const uint32_t oldw = __ldcg(p.dec + (j >> 5));
                if (bit != ((oldw >> (j & 31)) & 1u)) {
                    atomicXor(p.dec + (j >> 5), 1u << (j & 31));
                    for (int e = cp; e < cp + deg; ++e) {
                        const int c = p.ve_chk[e];
                        const uint32_t cb = 1u << (c & 31);
                        dn += (atomicXor(p.resid + (c >> 5), cb) & cb) ? -1 : 1;
                    }
                }
            }
            if (dn) atomicAdd(nnz, dn);
            grid_barrier(p.bar, gridDim.x, epoch);
            conv = *reinterpret_cast<volatile int *>(nnz) == 0;            // H*err mod 2 == syndrome (:180-181)
            ++iter;
            if (p.early_stop && conv) break;                               // :182-184
        }
        for (long long w = tid; w < p.NW; w += NT) p.err_words[b * p.NW + w] = __ldcg(p.dec + w);
        if (tid == 0) {
            p.conv[b] = conv ? 1 : 0;
            if (p.iters) p.iters[b] = iter;
            if (p.counters) {
                atomicAdd(&p.counters[0], 1ull);
                atomicAdd(&p.counters[1], conv ? 1ull : 0ull);
                atomicAdd(&p.counters[2], static_cast<unsigned long long>(iter));
            }
        }
        grid_barrier(p.bar, gridDim.x, epoch);                             // everyone has read nnz / dec of this syndrome
    }
    // leave the barrier counter at zero for the next launch (all CTAs have passed the last barrier: the slowest reader
    // only needed counter >= target, which stays true until this reset -- so reset by the last CTA to get here)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.bar + 1, 1u) == gridDim.x - 1) { p.bar[0] = 0u; p.bar[1] = 0u; p.work[0] = 0; p.work[1] = 0; }
    }
}

}  // inline namespace BP_VNS
}  // namespace bp
