// bp_global.cuh -- family GLOBAL: messages in HBM/L2, edge-major x syndrome-minor slabs.
//
// Same arithmetic and the same slot/refill scheme as family SMEM (bp_smem.cuh), for codes whose
// messages do not fit in shared memory (HGP-1600, regular n=100k, config C1's n=1000 single
// decode!).  Replaces /root/reference/src/decoders/belief_propagation.jl:121-188,220-231.
//
// Layout (per device)
//   NS resident syndrome slots, grouped in slabs of 32.  msg[slab][slot_edge][32] doubles:
//   lane = slot inside the slab, so a warp reads/writes 256 contiguous bytes per edge and every
//   global access of both passes is fully coalesced; node indices are warp-uniform.
//   syn/resid [slab][SW][32], errb [slab][NW][32] uint32, per-slot sid/iter/flags/nnz.
// One BP iteration = three launches: check pass, variable pass, finish (convergence test via
// nnz, output, refill of finished slots from a global queue counter).
#pragma once
#include "bp_math.cuh"

namespace bp {

struct GlobalParams {
    int s, n, E;
    int SW, NW;
    int max_iters, early_stop;
    int nslab;
    double p0;
    int regular_p0;
    long long B;
    const uint32_t *syn_words;
    uint32_t *err_words;
    uint8_t *conv;
    int32_t *iters;
    double *ratio;
    unsigned long long *counters;      // user counters (nullable)
    // graph tables
    const int *rowptr;                 // [s+1]
    const int *colptr;                 // [n+1]
    const int *ve_slot;                // [E]  CSC position -> check-major edge slot
    const int *ve_chk;                 // [E]  CSC position -> check index
    // state
    double *msg;                       // [nslab][E][32]
    uint32_t *syn, *resid, *errb;      // [nslab][SW|NW][32]
    long long *sid;                    // [NS]
    int *iter;                         // [NS]
    int *flags;                        // [NS] bit0 active, bit1 fresh
    int *nnz;                          // [NS]
    unsigned long long *queue;         // [0] next unclaimed syndrome, [1] syndromes finished
};

constexpr int kFlagActive = 1, kFlagFresh = 2;

template <bool BIG>
__global__ void __launch_bounds__(256) bp_global_check(const __grid_constant__ GlobalParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long total = static_cast<long long>(p.nslab) * p.s;
    for (long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
        const int slab = static_cast<int>(t / p.s);
        const int i = static_cast<int>(t - static_cast<long long>(slab) * p.s);
        const int fl = p.flags[slab * 32 + lane];
        if (!(fl & kFlagActive)) continue;
        const bool fresh = fl & kFlagFresh;
        const int rp = p.rowptr[i];
        const int deg = p.rowptr[i + 1] - rp;
        const bool neg = (p.syn[(static_cast<size_t>(slab) * p.SW + (i >> 5)) * 32 + lane] >> (i & 31)) & 1u;
        double *base = p.msg + (static_cast<size_t>(slab) * p.E + rp) * 32 + lane;
#define BP_CASE(D)                                                                   \
    {                                                                                \
        double m[D];                                                                 \
        if (fresh) { _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = p.p0; }    \
        else { _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = base[k * 32]; }  \
        check_update<D>(m, neg);                                                     \
        _Pragma("unroll") for (int k = 0; k < D; ++k) base[k * 32] = m[k];           \
    }
        BP_DEGREE_SWITCH(
            deg, BP_CASE, if (BIG) {
                check_update_big([&](int k) -> double & { return base[k * 32]; }, deg, neg, fresh, p.p0);
            })
#undef BP_CASE
    }
}

template <bool BIG>
__global__ void __launch_bounds__(256) bp_global_var(const __grid_constant__ GlobalParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long total = static_cast<long long>(p.nslab) * p.n;
    for (long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
        const int slab = static_cast<int>(t / p.n);
        const int j = static_cast<int>(t - static_cast<long long>(slab) * p.n);
        const int slot = slab * 32 + lane;
        if (!(p.flags[slot] & kFlagActive)) continue;
        const int cp = p.colptr[j];
        const int deg = p.colptr[j + 1] - cp;
        double *mb = p.msg + static_cast<size_t>(slab) * p.E * 32 + lane;
        double R = p.p0;
#define BP_CASE(D)                                                                   \
    {                                                                                \
        int v[D];                                                                    \
        double m[D];                                                                 \
        _Pragma("unroll") for (int k = 0; k < D; ++k) v[k] = p.ve_slot[cp + k];      \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = mb[static_cast<size_t>(v[k]) * 32]; \
        R = var_update<D>(m, p.p0, p.regular_p0 != 0);                                                  \
        _Pragma("unroll") for (int k = 0; k < D; ++k) mb[static_cast<size_t>(v[k]) * 32] = m[k]; \
    }
        BP_DEGREE_SWITCH(
            deg, BP_CASE, if (BIG) {
                R = var_update_big(
                    [&](int k) -> double & { return mb[static_cast<size_t>(p.ve_slot[cp + k]) * 32]; }, deg, p.p0);
            })
#undef BP_CASE
        if (p.ratio) p.ratio[p.sid[slot] * p.n + j] = R;
        const uint32_t e_new = (R >= 1.0) ? 1u : 0u;
        uint32_t *ew = p.errb + (static_cast<size_t>(slab) * p.NW + (j >> 5)) * 32 + lane;
        if (((*ew >> (j & 31)) & 1u) != e_new) {
            atomicXor(ew, 1u << (j & 31));
            int delta = 0;
            for (int k = 0; k < deg; ++k) {
                const int chk = p.ve_chk[cp + k];
                const uint32_t bit = 1u << (chk & 31);
                const uint32_t old =
                    atomicXor(p.resid + (static_cast<size_t>(slab) * p.SW + (chk >> 5)) * 32 + lane, bit);
                delta += (old & bit) ? -1 : 1;
            }
            if (delta) atomicAdd(p.nnz + slot, delta);
        }
    }
}

// One CTA per slab.  first = 1 on the launch that only fills the slots.
__global__ void __launch_bounds__(256) bp_global_finish(const __grid_constant__ GlobalParams p, int first)
{
    __shared__ long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int slab = blockIdx.x;
    const int slot = slab * 32 + lane;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // every warp evaluates the same slot state
    int fl = first ? 0 : p.flags[slot];
    const bool active = fl & kFlagActive;
    const int it = active ? p.iter[slot] + 1 : 0;
    const int cur_nnz = active ? p.nnz[slot] : 0;
    const bool conv = active && cur_nnz == 0;
    const bool done = first ? true : (active && ((p.early_stop && conv) || it >= p.max_iters));
    const long long old_sid = active ? p.sid[slot] : -1;
    const uint32_t done_mask = __ballot_sync(0xffffffffu, done);
    const int ndone = __popc(done_mask);

    if (threadIdx.x == 0) {
        s_base = ndone ? static_cast<long long>(atomicAdd(p.queue, static_cast<unsigned long long>(ndone))) : 0;
    }
    __syncthreads();
    long long new_sid = -1;
    if (done) {
        new_sid = s_base + __popc(done_mask & lt_mask);
        if (new_sid >= p.B) new_sid = -1;
    }
    // outputs of finished syndromes and refill, words spread over the CTA's warps
    if (!first) {
        for (int w = warp; w < p.NW; w += W) {
            if (done && old_sid >= 0)
                p.err_words[old_sid * p.NW + w] = p.errb[(static_cast<size_t>(slab) * p.NW + w) * 32 + lane];
        }
    }
    __syncthreads();    // errb fully read before it is cleared
    int cnt = 0;
    for (int w = warp; w < p.SW; w += W) {
        if (new_sid >= 0) {
            const uint32_t v = p.syn_words[new_sid * p.SW + w];
            p.syn[(static_cast<size_t>(slab) * p.SW + w) * 32 + lane] = v;
            p.resid[(static_cast<size_t>(slab) * p.SW + w) * 32 + lane] = v;
            cnt += __popc(v);
        }
    }
    for (int w = warp; w < p.NW; w += W)
        if (new_sid >= 0) p.errb[(static_cast<size_t>(slab) * p.NW + w) * 32 + lane] = 0u;
    // nnz of the new syndromes: sum the per-warp partial popcounts
    __shared__ int s_cnt[32];
    if (warp == 0) s_cnt[lane] = 0;
    __syncthreads();
    if (cnt) atomicAdd(&s_cnt[lane], cnt);
    __syncthreads();
    if (warp == 0) {
        if (done) {
            if (!first && old_sid >= 0) {
                p.conv[old_sid] = conv ? 1 : 0;
                if (p.iters) p.iters[old_sid] = it;
            }
            p.sid[slot] = new_sid;
            p.iter[slot] = 0;
            p.flags[slot] = (new_sid >= 0) ? (kFlagActive | kFlagFresh) : 0;
            p.nnz[slot] = s_cnt[lane];
        } else if (active) {
            p.iter[slot] = it;
            p.flags[slot] = kFlagActive;
        }
        // counters
        unsigned long long d = (!first && done) ? 1ull : 0ull, c = (!first && done && conv) ? 1ull : 0ull,
                           its = (!first && done) ? static_cast<unsigned long long>(it) : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            d += __shfl_xor_sync(0xffffffffu, d, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            its += __shfl_xor_sync(0xffffffffu, its, o);
        }
        if (lane == 0 && d) {
            atomicAdd(p.queue + 1, d);
            if (p.counters) {
                atomicAdd(p.counters + 0, d);
                atomicAdd(p.counters + 1, c);
                atomicAdd(p.counters + 2, its);
            }
        }
    }
}

}  // namespace bp
