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
//          incrementally: each warp holds the decisions of its variables as a bit field in a
//          register, resid = s xor H*e bit-packed in shared memory, nnz = popcount(resid).  A
//          variable whose decision flips XORs the bits of its checks in resid (shared-memory
//          atomics, lanes walk their own flips).  converged <=> nnz == 0.  A finished lane ORs
//          its set bits into the pre-zeroed packed output row.
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
    int uni_cdeg, uni_vdeg;   // common check / variable degree if the code is regular in it (1..12), else 0
    int SW, NW;               // uint32 words per packed syndrome / error row
    int max_iters;
    int early_stop;           // 1 = reference semantics
    double p0;                // per / (1 - per)
    int regular_p0;           // p0 is a positive normal double (no NaN clamp can fire on finite messages)
    long long B;
    const uint32_t *syn_words;    // [B][SW]
    uint32_t *err_words;          // [B][NW]
    uint8_t *conv;                // [B]
    int32_t *iters;               // [B] or null
    double *ratio;                // [B][n] or null
    unsigned long long *counters; // [4] or null
    const unsigned char *tables;  // global blob: rowptr u16[s+1] | colptr u16[n+1] | ve_off u32[E] | vflip u16[E]
    int tables_bytes;             // multiple of 16
    int off_colptr, off_ve, off_vflip;   // byte offsets inside the blob (rowptr at 0)
    // shared-memory carve-up (byte offsets from the dynamic smem base)
    int off_syn, off_resid, off_stage, off_nnz, off_tables, off_mbar;
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

// ---- explicit shared-window accesses: addresses are 32-bit shared-space offsets computed once,
// so the hot loops carry no generic->shared conversions.
template <int OFF = 0>
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF = 0>
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
template <int OFF = 0>
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
    return v;
}

template <int D, int K = 0>
__device__ __forceinline__ void load_row(double (&m)[D], uint32_t a)
{
    if constexpr (K < D) {
        m[K] = lds_f64<K * 256>(a);
        load_row<D, K + 1>(m, a);
    }
}
template <int D, int K = 0>
__device__ __forceinline__ void store_row(const double (&m)[D], uint32_t a)
{
    if constexpr (K < D) {
        sts_f64<K * 256>(a, m[K]);
        store_row<D, K + 1>(m, a);
    }
}
template <int D, int K = 0>
__device__ __forceinline__ void load_offsets(uint32_t (&v)[D], uint32_t a)
{
    if constexpr (K < D) {
        v[K] = lds_u32<K * 4>(a);
        load_offsets<D, K + 1>(v, a);
    }
}

// One check node of degree D whose D message slots start at shared address `a` (this lane's
// column, consecutive slots 256 B apart).  fresh lanes have not stored messages yet: they read
// the prior ratio instead (initialisation :127-131 without a store pass).
template <int D>
__device__ __forceinline__ void check_node(uint32_t a, bool neg, bool fresh, double p0)
{
    double m[D];
    load_row<D>(m, a);
    if (fresh) {
#pragma unroll
        for (int k = 0; k < D; ++k) m[k] = p0;
    }
    check_update<D>(m, neg);
    store_row<D>(m, a);
}

// One variable node of degree D; `vea` = shared address of its D slot offsets, `ml` = shared
// address of this lane's message column.  Returns the posterior ratio.
template <int D>
__device__ __forceinline__ double var_node(uint32_t ml, uint32_t vea, double p0, bool regular_p0)
{
    uint32_t v[D];
    double m[D];
    load_offsets<D>(v, vea);
#pragma unroll
    for (int k = 0; k < D; ++k) m[k] = lds_f64(ml + v[k]);
    const double R = var_update<D>(m, p0, regular_p0);
#pragma unroll
    for (int k = 0; k < D; ++k) sts_f64(ml + v[k], m[k]);
    return R;
}

// MAXT / MINB: launch bounds (instantiated for 2 CTAs/SM with 256/320/384 threads and 1 CTA/SM with 512)
template <bool BIG, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bp_smem_kernel(const __grid_constant__ SmemParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *syn = reinterpret_cast<uint32_t *>(smem + p.off_syn);
    uint32_t *resid = reinterpret_cast<uint32_t *>(smem + p.off_resid);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem + p.off_stage);
    int *nnz = reinterpret_cast<int *>(smem + p.off_nnz);          // [2][32]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + p.off_mbar);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // shared-window addresses used by the hot loops
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ml = sbase + lane * 8;                          // this lane's message column
    const uint32_t syn_a = sbase + p.off_syn + lane * 4;
    const uint32_t rowptr_a = sbase + p.off_tables;
    const uint32_t colptr_a = sbase + p.off_tables + p.off_colptr;
    const uint32_t ve_a = sbase + p.off_tables + p.off_ve;         // u32 byte offset of each edge's slot row
    const uint32_t vflip_a = sbase + p.off_tables + p.off_vflip;   // u16 residual word offset | bit of each edge's check

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
    // Hard decisions of the variables this warp owns (j = warp + i*W  <->  bit i), one bit set per
    // variable currently decided 1.  Host guarantees ceil(n / W) <= 64.
    unsigned long long ebits = 0;
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
            ebits = 0;                                                      // err .= 0 (reset!, :89)
            if (warp == 0 && active) {
                int cnt = 0;
                for (int w = 0; w < p.SW; ++w) {
                    const uint32_t v = stage[w * 32 + rank];
                    syn[w * 32 + lane] = v;
                    resid[w * 32 + lane] = v;
                    cnt += __popc(v);
                }
                nnz[nnz_buf * 32 + lane] = cnt;
            }
        }
        q_head += __popc(mask);
        if (warp == 0) { __syncwarp(); prefetch(q_head); }
    };

    refill(0xffffffffu, par);
    __syncthreads();

    const double p0 = p.p0;
    const bool regular_p0 = p.regular_p0;
    while (__ballot_sync(0xffffffffu, active) != 0u) {
        // ------------------------------------------------------------------ check pass (:135-150)
        // warp w owns checks w, w+W, ...; syndrome bit of check i = bit i%32 of syn[i/32][lane]
        if (active) {
            if (p.uni_cdeg) {
                // every check has the same degree: slots of check i start at i*D, no table reads
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        uint32_t a = ml + warp * (D * 256);                                                      \
        for (int i = warp; i < p.s; i += W, a += W * (D * 256)) {                                \
            const bool neg = (lds_u32(syn_a + (i >> 5) * 128) >> (i & 31)) & 1u;                 \
            check_node<D>(a, neg, fresh, p0);                                                    \
        }                                                                                        \
    }
                BP_DEGREE_SWITCH(p.uni_cdeg, BP_CASE, ;)
#undef BP_CASE
            } else {
                for (int i = warp; i < p.s; i += W) {
                    const int rp = lds_u16(rowptr_a + 2 * i);
                    const int deg = lds_u16(rowptr_a + 2 * i + 2) - rp;
                    const bool neg = (lds_u32(syn_a + (i >> 5) * 128) >> (i & 31)) & 1u;
                    const uint32_t a = ml + rp * 256;
#define BP_CASE(D) check_node<D>(a, neg, fresh, p0)
                    BP_DEGREE_SWITCH(
                        deg, BP_CASE, if (BIG) {
                            double *base = reinterpret_cast<double *>(smem + lane * 8 + rp * 256);
                            check_update_big([&](int k) -> double & { return base[k * 32]; }, deg, neg, fresh, p0);
                        })
#undef BP_CASE
                }
            }
        }
        __syncthreads();
        // --------------------------------------------------------------- variable pass (:152-178)
        // warp w owns variables w, w+W, ...; decision of its i-th variable = bit i of newbits
        if (active) {
            unsigned long long newbits = 0;
            if (p.uni_vdeg) {
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        uint32_t vea = ve_a + warp * (D * 4);                                                    \
        int i = 0;                                                                               \
        for (int j = warp; j < p.n; j += W, ++i, vea += W * (D * 4)) {                           \
            const double R = var_node<D>(ml, vea, p0, regular_p0);                               \
            if (p.ratio) p.ratio[sid * p.n + j] = R;                                             \
            newbits |= static_cast<unsigned long long>((R >= 1.0) ? 1u : 0u) << i;               \
        }                                                                                        \
    }
                BP_DEGREE_SWITCH(p.uni_vdeg, BP_CASE, ;)
#undef BP_CASE
            } else {
                int i = 0;
                for (int j = warp; j < p.n; j += W, ++i) {
                    const int cp = lds_u16(colptr_a + 2 * j);
                    const int deg = lds_u16(colptr_a + 2 * j + 2) - cp;
                    const uint32_t vea = ve_a + cp * 4;
                    double R = p0;                                                    // degree 0: prior only
#define BP_CASE(D) R = var_node<D>(ml, vea, p0, regular_p0)
                    BP_DEGREE_SWITCH(
                        deg, BP_CASE, if (BIG) {
                            const uint32_t *ve = reinterpret_cast<const uint32_t *>(smem + p.off_tables + p.off_ve) + cp;
                            unsigned char *msgb = smem + lane * 8;
                            R = var_update_big([&](int k) -> double & { return *reinterpret_cast<double *>(msgb + ve[k]); },
                                               deg, p0);
                        })
#undef BP_CASE
                    if (p.ratio) p.ratio[sid * p.n + j] = R;
                    newbits |= static_cast<unsigned long long>((R >= 1.0) ? 1u : 0u) << i;   // :164-168 (tie -> 1)
                }
            }
            // Only variables whose decision flipped touch the residual syndrome s xor H*e
            // (syndrome re-check :180-181, kept incrementally; lanes walk their own flips).
            unsigned long long flips = ebits ^ newbits;
            ebits = newbits;
            if (flips) {
                int delta = 0;
                do {
                    const int j = warp + (__ffsll(static_cast<long long>(flips)) - 1) * W;
                    flips &= flips - 1;
                    const int e1 = lds_u16(colptr_a + 2 * j + 2);
                    for (int e = lds_u16(colptr_a + 2 * j); e < e1; ++e) {
                        const uint32_t ent = lds_u16(vflip_a + 2 * e);        // (check/32)*128 + check%32
                        const uint32_t old = atomicXor(reinterpret_cast<uint32_t *>(smem + p.off_resid + lane * 4 + (ent & 0xff80u)),
                                                       1u << (ent & 31u));
                        delta += 1 - 2 * static_cast<int>((old >> (ent & 31u)) & 1u);
                    }
                } while (flips);
                if (delta) atomicAdd(nnz + par * 32 + lane, delta);
            }
        }
        __syncthreads();
        // ---------------------------------------- syndrome re-check, early stop, refill (:180-184)
        const int cur_nnz = nnz[par * 32 + lane];
        if (active) { ++iter; fresh = false; }
        const bool conv = active && cur_nnz == 0;
        const bool done = active && ((p.early_stop && conv) || iter >= p.max_iters);
        const uint32_t done_mask = __ballot_sync(0xffffffffu, done);
        if (done) {
            // errors[:, sid] = guess (:227): every warp ORs the set bits it owns into the
            // pre-zeroed packed row
            unsigned long long b = ebits;
            while (b) {
                const int j = warp + (__ffsll(static_cast<long long>(b)) - 1) * W;
                b &= b - 1;
                atomicOr(p.err_words + sid * p.NW + (j >> 5), 1u << (j & 31));
            }
        }
        if (warp == 0) {
            if (done) {
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
