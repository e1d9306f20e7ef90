// bp_filter.cuh -- the first BP iteration of every syndrome without messages ("first-iteration filter").
//
// In iteration 1 of decode! (/root/reference/src/decoders/belief_propagation.jl:127-184) every bit->check message is the
// prior p/(1-p), so a check's outputs depend only on its degree, the position inside the check and the check's syndrome
// bit, and a variable's posterior ratio -- hence its hard decision -- only on the syndrome bits of its own checks.  For a
// variable of degree d that is a truth table of 2^d bits.  first_iter_tables_kernel computes those tables with the
// kernels' own node updates (check_update<D> on all-prior inputs, var_update<D> on the resulting constants: the very
// instruction sequences a fresh lane of the decoding kernels executes, so the bits are the reference's by
// construction); first_iter_filter_kernel then evaluates iteration 1 for a whole batch with integer instructions only:
// decisions e1 = table(syndrome), residual s xor H*e1, and
//   * converged (:182-184 in iteration 1): outputs written here (errors = e1, converged = 1, 1 iteration),
//   * otherwise: the syndrome's index is appended to a work list that the decoding kernel takes as its queue (it starts
//     those from scratch, iteration 1 included).
// At per = 0.03 on the gross code 42 % of the syndromes end here; at 0.01, 90 %.  Variable degrees up to
// kFilterMaxVarDeg, check degrees up to kMaxRegDegree, early stop on, posterior ratios not requested for every
// iteration: otherwise the filter is simply not used.
#pragma once
#include "bp_math.cuh"

namespace bp {

constexpr int kFilterMaxVarDeg = 5;

struct FilterVar {              // 16 bytes per variable
    uint16_t chk[kFilterMaxVarDeg];   // its checks, ascending
    uint8_t deg, pad;
    uint32_t tt;                      // decision for every pattern of its checks' syndrome bits (bit k of the index <-> chk[k])
};

struct FilterParams {
    int s, n, SW, NW;
    long long B;
    const FilterVar *vars;
    const uint32_t *syn_words;
    uint32_t *err_words;          // rows must be zero on entry (finished lanes OR bits in)
    int out_bits;                 // err_words is a bit stream (bit b*n + j, Julia BitMatrix) instead of rows of NW words
    uint8_t *conv;
    int32_t *iters;
    int *list, *list_count;       // work list of the syndromes that need more than one iteration
    unsigned long long *counters; // [0] decoded, [1] converged, [2] iterations, [3] finished by this filter (added for the syndromes finished here)
};

constexpr int kFilterThreads = 128;

#ifdef BP_FILTER_KERNEL      // (ldpcb200.cu only: the kernel is variant-independent)
// 32 x 32 bit-matrix transpose across a warp: lane L passes row L, lane k receives column k (bit L of the result = bit k
// of lane L's word).  Five exchange steps of block swaps.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane)
{
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
    }
    return x;
}

// Bit-sliced: a warp takes 32 syndromes at a time.  (A) the syndrome words are transposed so that word T[i] holds bit i of
// all 32 syndromes; (B) the lanes share out the variables: a variable's decisions for the 32 syndromes are ONE evaluation
// of its truth table on the words of its checks, and they are folded into the residual words R[i] = T[i] xor (H e)_i;
// (C) OR over the checks gives the mask of syndromes that did not converge; (D) the decision words are transposed back
// into packed error rows for the converged ones, the others are appended to the work list (one atomic per warp).
// Shared memory per warp: T[SW*32] | R[SW*32] | E[NW*32] words.
constexpr int kFilterWarps = kFilterThreads / 32;
__global__ void __launch_bounds__(kFilterThreads) first_iter_filter_kernel(const FilterParams p)
{
    extern __shared__ uint32_t fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sp = p.SW * 32, np_ = p.NW * 32;
    uint32_t *T = fsm + warp * (2 * sp + np_);
    uint32_t *R = T + sp;
    uint32_t *E = R + sp;
    unsigned long long n_conv = 0;
    const long long nblk = (p.B + 31) >> 5;
    for (long long blk = static_cast<long long>(blockIdx.x) * kFilterWarps + warp; blk < nblk; blk += static_cast<long long>(gridDim.x) * kFilterWarps) {
        const long long b = (blk << 5) + lane;
        const bool live = b < p.B;
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        // (A) transpose: lane k ends up with T[32 w + k]
        for (int w = 0; w < p.SW; ++w) {
            const uint32_t x = warp_transpose32(live ? p.syn_words[b * p.SW + w] : 0u, lane);
            T[w * 32 + lane] = x;
            R[w * 32 + lane] = x;
        }
        for (int w = 0; w < p.NW; ++w) E[w * 32 + lane] = 0u;
        __syncwarp();
        // (B) decisions of every variable for the 32 syndromes at once
        for (int j = lane; j < p.n; j += 32) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p.vars) + j);
            const uint32_t tt = raw.w;
            const int deg = (raw.z >> 16) & 0xff;
            uint32_t c[kFilterMaxVarDeg] = {raw.x & 0xffffu, raw.x >> 16, raw.y & 0xffffu, raw.y >> 16, raw.z & 0xffffu};
            uint32_t t[kFilterMaxVarDeg];
#pragma unroll
            for (int k = 0; k < kFilterMaxVarDeg; ++k) t[k] = k < deg ? T[c[k]] : 0u;
            uint32_t f = 0;
            for (uint32_t mt = 0; mt < (1u << deg); ++mt)
                if ((tt >> mt) & 1u) {
                    uint32_t a = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < kFilterMaxVarDeg; ++k)
                        if (k < deg) a &= ((mt >> k) & 1u) ? t[k] : ~t[k];
                    f |= a;
                }
            E[j] = f;
            if (f) {
#pragma unroll
                for (int k = 0; k < kFilterMaxVarDeg; ++k)
                    if (k < deg) atomicXor(&R[c[k]], f);
            }
        }
        __syncwarp();
        // (C) syndromes with an unsatisfied check
        uint32_t un = 0;
        for (int i = lane; i < p.s; i += 32) un |= R[i];
        un = __reduce_or_sync(0xffffffffu, un) & live_mask;
        const uint32_t cv = ~un & live_mask;
        // (D) outputs
        for (int w = 0; w < p.NW; ++w) {
            const uint32_t row = warp_transpose32(E[w * 32 + lane], lane);      // decisions 32 w .. 32 w + 31 of syndrome `lane`
            if (((cv >> lane) & 1u) && row) {
                if (p.out_bits) {                                  // neighbouring syndromes share words: OR the two halves in
                    const unsigned long long o = static_cast<unsigned long long>(b) * static_cast<unsigned long long>(p.n) + 32ull * w;
                    const int lo = static_cast<int>(o & 31ull);
                    atomicOr(p.err_words + (o >> 5), row << lo);
                    if (lo && (row >> (32 - lo))) atomicOr(p.err_words + (o >> 5) + 1, row >> (32 - lo));
                } else {
                    p.err_words[b * p.NW + w] = row;
                }
            }
        }
        if ((cv >> lane) & 1u) {
            p.conv[b] = 1;
            if (p.iters) p.iters[b] = 1;
        }
        if (un) {
            int base = 0;
            if (lane == 0) base = atomicAdd(p.list_count, __popc(un));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((un >> lane) & 1u) p.list[base + __popc(un & ((1u << lane) - 1u))] = static_cast<int>(b);
        }
        n_conv += __popc(cv);
        __syncwarp();
    }
    if (p.counters && lane == 0 && n_conv) {
        atomicAdd(p.counters + 0, n_conv);
        atomicAdd(p.counters + 1, n_conv);
        atomicAdd(p.counters + 2, n_conv);
        atomicAdd(p.counters + 3, n_conv);      // LDPCB200_CTR_FILTERED
    }
}
#endif  // BP_FILTER_KERNEL

struct FilterSetup {
    int n;
    double p0, check_aux;
    int regular_p0;
    const int *colptr;          // [n+1]
    const uint8_t *e_deg;       // [E] degree of the check of each CSC edge
    const uint8_t *e_pos;       // [E] position of the variable inside that check
    double *ctab;               // [kMaxRegDegree + 1][2][kMaxRegDegree] check outputs on all-prior inputs: [degree][syndrome bit][position]
    FilterVar *vars;            // tt is written here
};

#ifdef BP_FILTER_SETUP       // (the per-variant translation units: the tables come from that variant's node updates)
inline namespace BP_VNS {

// grid 1 x (2 * kMaxRegDegree) threads: thread (d-1)*2 + neg
__global__ void first_iter_check_table_kernel(const FilterSetup q)
{
    const int t = threadIdx.x;
    if (t >= 2 * kMaxRegDegree) return;
    const int deg = t / 2 + 1;
    const bool neg = t & 1;
    double *out = q.ctab + (deg * 2 + (neg ? 1 : 0)) * kMaxRegDegree;
#define BP_CASE(D)                                                     \
    {                                                                  \
        double m[D];                                                   \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = q.p0;     \
        check_update<D>(m, neg, q.check_aux);                          \
        _Pragma("unroll") for (int k = 0; k < D; ++k) out[k] = m[k];   \
    }
    BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
}

// one thread per variable: its 2^deg decisions
__global__ void first_iter_truth_table_kernel(const FilterSetup q)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= q.n) return;
    const int e0 = q.colptr[j], deg = q.colptr[j + 1] - e0;
    uint32_t tt = 0;
    for (uint32_t idx = 0; idx < (1u << deg); ++idx) {
        double R = q.p0;
#define BP_CASE(D)                                                                                              \
    {                                                                                                           \
        double m[D];                                                                                            \
        _Pragma("unroll") for (int k = 0; k < D; ++k)                                                           \
            m[k] = q.ctab[(q.e_deg[e0 + k] * 2 + ((idx >> k) & 1u)) * kMaxRegDegree + q.e_pos[e0 + k]];         \
        R = var_update<D>(m, q.p0, q.regular_p0 != 0);                                                          \
    }
        switch (deg) {
            case 1: BP_CASE(1) break;
            case 2: BP_CASE(2) break;
            case 3: BP_CASE(3) break;
            case 4: BP_CASE(4) break;
            case 5: BP_CASE(5) break;
            default: break;                  // degree 0: the prior alone
        }
#undef BP_CASE
        if (decide(R)) tt |= 1u << idx;
    }
    q.vars[j].tt = tt;
}

}  // inline namespace BP_VNS
#endif  // BP_FILTER_SETUP
}  // namespace bp
