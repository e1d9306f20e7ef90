// bp_smem.cuh -- family SMEM: one persistent kernel runs every BP iteration of every syndrome.
//
// Replaces the whole of decode!/batchdecode! (/root/reference/src/decoders/belief_propagation.jl
// :121-188, :220-231) for codes whose per-syndrome message array (E doubles) times 32 syndromes
// fits in shared memory (surface d<=15, [[144,12,12]] gross code, ...).
//
// Mapping
//   CTA  = 32 "lanes" (syndrome slots) x W warps.  Lane l of EVERY warp works on the syndrome
//          currently held in slot l; warp w owns checks w, w+W, ... in the check pass and
//          variables w, w+W, ... in the variable pass.  All node indices are therefore
//          warp-uniform (broadcast table reads, no divergence) and every message access is
//          msg[slot_edge*32 + lane]: 256 contiguous bytes per warp, bank-conflict free.
//   Messages: ONE in-place array, check-major edge order (a check's edges are contiguous);
//          the check pass turns bit->check ratios into check->bit ratios in place, the variable
//          pass turns them back.
//   Hard decisions and the syndrome re-check (belief_propagation.jl:164-168,180-184) are kept
//          incrementally: errb = bit-packed current decision, resid = s xor H*e bit-packed,
//          nnz = popcount(resid).  A variable whose decision flips XORs its bit in errb and the
//          bits of its checks in resid (shared-memory atomics, rare).  converged <=> nnz == 0.
//   Early termination / compaction: every lane has its own iteration counter.  A lane whose
//          syndrome converged (or hit max_iters) writes its outputs and immediately takes the
//          next syndrome of the CTA's queue, so the FP64 pipe never idles on finished
//          syndromes.  Fresh lanes read the prior p/(1-p) instead of stored messages
//          (initialisation :127-131 without a store pass).
//   Edge tables are fetched into shared memory with one TMA bulk copy (cp.async.bulk +
//          mbarrier); the next 32 queued syndromes are prefetched with cp.async.
#pragma once
#include "bp_math.cuh"

namespace bp {

struct SmemParams {
    int s, n, E;
    int SW, NW;               // uint32 words per packed syndrome / error row
    int max_iters;
    int early_stop;           // 1 = reference semantics
    double p0;                // per / (1 - per)
    long long B;
    const uint32_t *syn_words;    // [B][SW]
    uint32_t *err_words;          // [B][NW]
    uint8_t *conv;                // [B]
    int32_t *iters;               // [B] or null
    double *ratio;                // [B][n] or null
    unsigned long long *counters; // [4] or null
    const unsigned char *tables;  // global blob: rowptr u16[s+1] | colptr u16[n+1] | ve_off u32[E] | ve_chk u16[E]
    int tables_bytes;             // multiple of 16
    int off_colptr, off_ve, off_vchk;   // byte offsets inside the blob (rowptr at 0)
    // shared-memory carve-up (byte offsets from the dynamic smem base)
    int off_syn, off_resid, off_errb, off_stage, off_nnz, off_tables, off_mbar;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_tables(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(b)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t b = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// MAXT / MINB: launch bounds (two instantiations: 2 CTAs/SM with up to 384 threads, 1 CTA/SM with 512)
template <bool BIG, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bp_smem_kernel(const __grid_constant__ SmemParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *syn = reinterpret_cast<uint32_t *>(smem + p.off_syn);
    uint32_t *resid = reinterpret_cast<uint32_t *>(smem + p.off_resid);
    uint32_t *errb = reinterpret_cast<uint32_t *>(smem + p.off_errb);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem + p.off_stage);
    int *nnz = reinterpret_cast<int *>(smem + p.off_nnz);          // [2][32]
    const uint16_t *rowptr = reinterpret_cast<const uint16_t *>(smem + p.off_tables);
    const uint16_t *colptr = reinterpret_cast<const uint16_t *>(smem + p.off_tables + p.off_colptr);
    const uint32_t *ve = reinterpret_cast<const uint32_t *>(smem + p.off_tables + p.off_ve);       // byte offset of the edge's slot row
    const uint16_t *vchk = reinterpret_cast<const uint16_t *>(smem + p.off_tables + p.off_vchk);   // check of the edge
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + p.off_mbar);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned char *msgb = smem + lane * 8;         // this lane's column of the message array

    if (threadIdx.x == 0)
        tma_load_tables(smem + p.off_tables, p.tables, static_cast<uint32_t>(p.tables_bytes), mbar);

    // This CTA's queue: 32-syndrome chunks c, c+G, c+2G, ... of the batch.
    const long long G = gridDim.x, c = blockIdx.x;
    const long long nchunks = (p.B + 31) >> 5;
    const long long my_chunks = (c < nchunks) ? (nchunks - c + G - 1) / G : 0;
    const long long Q = my_chunks << 5;
    auto sid_of = [&](long long q) -> long long {
        const long long sid = (((q >> 5) * G + c) << 5) + (q & 31);
        return (q < Q && sid < p.B) ? sid : -1;
    };
    auto prefetch = [&](long long q_head) {      // warp 0: stage[w][r] <- syndrome words of entry q_head + r
        const long long sid = sid_of(q_head + lane);
        if (sid >= 0)
            for (int w = 0; w < p.SW; ++w) cp_async4(&stage[w * 32 + lane], p.syn_words + sid * p.SW + w);
    };

    long long q_head = 0;
    long long sid = -1;
    int iter = 0;
    bool active = false, fresh = false;
    int par = 0;                                  // which nnz buffer the coming iteration updates
    unsigned long long n_done = 0, n_conv = 0, n_iters = 0;   // warp 0 only

    if (warp == 0) prefetch(0);
    __syncthreads();
    mbar_wait(mbar, 0);                           // tables have landed

    // Lanes in `mask` take the next queue entries.  Executed identically by every warp
    // (register state is replicated); warp 0 additionally moves the staged syndrome in.
    auto refill = [&](uint32_t mask, int nnz_buf) {
        if (warp == 0) { cp_async_wait_all(); __syncwarp(); }
        if ((mask >> lane) & 1u) {
            const int rank = __popc(mask & lt_mask);
            sid = sid_of(q_head + rank);
            active = sid >= 0;
            fresh = active;
            iter = 0;
            if (warp == 0 && active) {
                int cnt = 0;
                for (int w = 0; w < p.SW; ++w) {
                    const uint32_t v = stage[w * 32 + rank];
                    syn[w * 32 + lane] = v;
                    resid[w * 32 + lane] = v;
                    cnt += __popc(v);
                }
                nnz[nnz_buf * 32 + lane] = cnt;
                for (int w = 0; w < p.NW; ++w) errb[w * 32 + lane] = 0u;   // err .= 0 (reset!, :89)
            }
        }
        q_head += __popc(mask);
        if (warp == 0) { __syncwarp(); prefetch(q_head); }
    };

    refill(0xffffffffu, par);
    __syncthreads();

    while (__ballot_sync(0xffffffffu, active) != 0u) {
        // ------------------------------------------------------------------ check pass (:135-150)
        if (active) {
            for (int i = warp; i < p.s; i += W) {
                const int rp = rowptr[i];
                const int deg = rowptr[i + 1] - rp;
                const bool neg = (syn[(i >> 5) * 32 + lane] >> (i & 31)) & 1u;
                double *base = reinterpret_cast<double *>(msgb + rp * 256);
#define BP_CASE(D)                                                                   \
    {                                                                                \
        double m[D];                                                                 \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = base[k * 32];           \
        if (fresh) { _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = p.p0; }    \
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
        __syncthreads();
        // --------------------------------------------------------------- variable pass (:152-178)
        if (active) {
            int *nz = nnz + par * 32 + lane;
            for (int j = warp; j < p.n; j += W) {
                const int cp = colptr[j];
                const int deg = colptr[j + 1] - cp;
                double R = p.p0;
#define BP_CASE(D)                                                                   \
    {                                                                                \
        uint32_t v[D];                                                               \
        double m[D];                                                                 \
        _Pragma("unroll") for (int k = 0; k < D; ++k) v[k] = ve[cp + k];             \
        _Pragma("unroll") for (int k = 0; k < D; ++k) m[k] = *reinterpret_cast<double *>(msgb + v[k]); \
        R = var_update<D>(m, p.p0);                                                  \
        _Pragma("unroll") for (int k = 0; k < D; ++k) *reinterpret_cast<double *>(msgb + v[k]) = m[k]; \
    }
                BP_DEGREE_SWITCH(
                    deg, BP_CASE, if (BIG) {
                        R = var_update_big(
                            [&](int k) -> double & { return *reinterpret_cast<double *>(msgb + ve[cp + k]); }, deg, p.p0);
                    })
#undef BP_CASE
                if (p.ratio) p.ratio[sid * p.n + j] = R;
                const uint32_t e_new = (R >= 1.0) ? 1u : 0u;                          // :164-168 (tie -> 1)
                uint32_t *ew = errb + (j >> 5) * 32 + lane;
                if (((*ew >> (j & 31)) & 1u) != e_new) {
                    atomicXor(ew, 1u << (j & 31));
                    int delta = 0;
                    for (int k = 0; k < deg; ++k) {
                        const uint32_t chk = vchk[cp + k];
                        const uint32_t bit = 1u << (chk & 31);
                        const uint32_t old = atomicXor(resid + (chk >> 5) * 32 + lane, bit);
                        delta += (old & bit) ? -1 : 1;
                    }
                    if (delta) atomicAdd(nz, delta);
                }
            }
        }
        __syncthreads();
        // ---------------------------------------- syndrome re-check, early stop, refill (:180-184)
        const int cur_nnz = nnz[par * 32 + lane];
        if (active) { ++iter; fresh = false; }
        const bool conv = active && cur_nnz == 0;
        const bool done = active && ((p.early_stop && conv) || iter >= p.max_iters);
        const uint32_t done_mask = __ballot_sync(0xffffffffu, done);
        if (warp == 0) {
            if (done) {
                for (int w = 0; w < p.NW; ++w) p.err_words[sid * p.NW + w] = errb[w * 32 + lane];
                p.conv[sid] = conv ? 1 : 0;
                if (p.iters) p.iters[sid] = iter;
                n_done += 1; n_conv += conv ? 1 : 0; n_iters += iter;
            } else {
                nnz[(par ^ 1) * 32 + lane] = cur_nnz;          // carry over to the other buffer
            }
        }
        par ^= 1;
        if (done_mask) refill(done_mask, par);
        __syncthreads();
    }

    if (warp == 0 && p.counters) {
        for (int o = 16; o > 0; o >>= 1) {
            n_done += __shfl_xor_sync(0xffffffffu, n_done, o);
            n_conv += __shfl_xor_sync(0xffffffffu, n_conv, o);
            n_iters += __shfl_xor_sync(0xffffffffu, n_iters, o);
        }
        if (lane == 0) {
            atomicAdd(p.counters + 0, n_done);
            atomicAdd(p.counters + 1, n_conv);
            atomicAdd(p.counters + 2, n_iters);
        }
    }
}

}  // namespace bp
