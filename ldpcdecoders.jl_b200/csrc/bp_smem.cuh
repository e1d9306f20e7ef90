// bp_smem.cuh -- the shared-memory-resident persistent BP kernel (family SMEM, round 2).
//
// Same mapping, arithmetic and results as bp_persistent_kernel<0, ...> (bp_kernel.cuh): CTA = 32
// syndrome lanes x W warps, messages of 32 syndromes resident in shared memory for all iterations,
// decode!/batchdecode! of /root/reference/src/decoders/belief_propagation.jl:121-188,220-231.
// What differs is everything AROUND the node updates, which ncu showed to be a third of all issued
// instructions of the round-1 kernel (profiles/r1_persistent_kernel_c3_mode0_ncu_full.txt):
//   * two block barriers per iteration instead of three: an entering syndrome is read from its staging slot
//     during its first check pass (per-lane syndrome row address), so the copy into the lane's own rows
//     needs no barrier of its own;
//   * the syndrome re-check (:180-184) is `residual == 0`, the residual s xor H*e being updated by
//     NON-returning shared-memory atomics of the flipped variables; no per-lane count of unsatisfied
//     checks, no returning atomics, no count exchange.  The residual of a lane is double-buffered per lane
//     (an entering syndrome takes the other buffer) so that a warp still testing the leaving syndrome never
//     sees the entering one;
//   * 32-bit queue arithmetic (one launch handles fewer than 2^31 syndromes; the host splits larger batches),
//     32-bit decision fields when a warp owns at most 32 variables;
//   * degree segments only (the degree switch runs once per segment), no per-node paths, no decision fields
//     in memory, no local-memory degrees: codes outside that envelope stay on the round-1 kernel.
#pragma once
#include <type_traits>

#include "bp_kernel.cuh"

namespace bp {
inline namespace BP_VNS {

__device__ __forceinline__ uint4 lds_u128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// NV variables of degree D whose NV*D slot offsets are consecutive in the table at `ta` (16-byte aligned when NV == 4):
// loads, products, one NaN test for all of them (bp_math.cuh: var_products), stores.  R[] receives the posterior ratios.
template <int D, int NV>
__device__ __forceinline__ void var_group(uint32_t ml, uint32_t ta, double p0, bool regular_p0, double (&R)[NV])
{
    uint32_t off[NV * D];
    if constexpr (NV == 4) {
#pragma unroll
        for (int q = 0; q < D; ++q) {
            const uint4 v = lds_u128(ta + 16 * q);
            off[4 * q] = v.x; off[4 * q + 1] = v.y; off[4 * q + 2] = v.z; off[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NV * D; ++q) off[q] = lds_u32(ta + 4 * q);
    }
    double m[NV][D], o[NV][D];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < D; ++k) m[v][k] = ld_msg(ml + off[v * D + k]);
    uint32_t flag = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) flag = max(flag, var_products<D>(m[v], p0, o[v], R[v]));
    if (var_products_suspect(flag) || !regular_p0) {       // rare: some running product is NaN -- the clamped sequence
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            R[v] = var_update_clamped<D>(m[v], p0);
#pragma unroll
            for (int k = 0; k < D; ++k) o[v][k] = m[v][k];
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < D; ++k) st_msg(ml + off[v * D + k], o[v][k]);
}

template <int MAXT, int MINB, bool EB64, bool PROF = false, bool DUAL = false>
__global__ void __launch_bounds__(MAXT, MINB) bp_smem_kernel(const __grid_constant__ KernelParams p)
{
    using ebits_t = typename std::conditional<EB64, unsigned long long, uint32_t>::type;
    extern __shared__ __align__(128) unsigned char smem[];
    // DUAL: one CTA per SM carries TWO independent groups of 32 syndromes ("teams" of W warps, each with its own
    // messages, state, queue and block barrier) that take turns in the FP64-bound check pass: a team enters its check
    // pass only when the other one has left its own, so that one team's check pass runs against the other's
    // variable pass / bookkeeping (which leave the FP64 pipe idle) instead of against its check pass.  Two independent
    // CTAs per SM drift into exactly that contention (measured: both in the check pass, then both out of it).
    const int lane = threadIdx.x & 31;
    const int W = DUAL ? (blockDim.x >> 6) : (blockDim.x >> 5);   // warps per team
    const int team = DUAL ? static_cast<int>(threadIdx.x >> 5) / W : 0;
    const int warp = static_cast<int>(threadIdx.x >> 5) - team * W;   // warp index inside the team
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t goff = DUAL ? static_cast<uint32_t>(team * p.group_stride) : 0u;   // this team's group of arrays
    const uint32_t gbase = sbase + goff;
    const uint32_t ml = gbase + lane * 8;                         // this lane's message column
    const uint32_t syn_own = gbase + p.off_syn + lane * 4;        // this lane's syndrome words, [word][32] layout
    const uint32_t res0 = gbase + p.off_resid + lane * 4;         // residual buffers 0 / 1 of this lane
    const uint32_t res_sum = 2 * res0 + p.SW * 128;               // res0 + res1
    const uint32_t stage_a = gbase + p.off_stage;                 // [2][SW][32] staged syndromes of the queue window
    const uint32_t sidq_a = gbase + p.off_sidq;                   // [2][32] their indices in the batch (-1: past the end)
    auto team_sync = [&]() {                                      // block barrier of this team
        if constexpr (DUAL) asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(W * 32) : "memory");
        else __syncthreads();
    };
    // The check-pass token (DUAL).  Hardware barrier 3 + t is waited on by team t and arrived at by the other team;
    // exactly one token exists (team 1 hands it to team 0 in the prologue) and a team only ever passes it on after
    // having received it, so arrivals can never pile up on a barrier nobody waits on.  A team that has drained its
    // queue takes one last (empty) turn, raises team_done and passes the token for good; the other team learns of it
    // through `other_done`, which one thread samples and the team reads after its next block barrier, so that all
    // warps of a team always agree on whether they take part in the hand-over.
    volatile int *team_done = reinterpret_cast<volatile int *>(smem + p.off_mbar + 8);   // [2] raised by a finished team
    volatile int *team_seen = team_done + 2;                                             // [2] team t's sample of team_done[t^1]
    bool other_done = false;
    auto token_wait = [&]() {
        if constexpr (DUAL) {
            if (!other_done) asm volatile("bar.sync %0, %1;" ::"r"(3 + team), "r"(2 * W * 32) : "memory");
        }
    };
    auto token_pass = [&]() {
        if constexpr (DUAL) {
            if (!other_done) asm volatile("bar.arrive %0, %1;" ::"r"(3 + (team ^ 1)), "r"(2 * W * 32) : "memory");
        }
    };
    const uint32_t colptr_a = sbase + p.off_tables + p.off_colptr;
    const uint32_t ve_a = sbase + p.off_tables + p.off_ve;        // u32 byte offset of each edge's slot row
    const uint32_t vflip_a = sbase + p.off_tables + p.off_vflip;  // u16 residual word offset | bit of each edge's check
    const uint32_t corig_a = sbase + p.off_tables + p.off_corig;
    const uint32_t vorig_a = sbase + p.off_tables + p.off_vorig;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + p.off_mbar);

    // contiguous variable ownership (uniform variable degree, caller's variable order): bit i of a warp's decision
    // field is variable vbase + i; otherwise variables are dealt round-robin (bit i <-> warp + i*W)
    const bool cv = p.cv_cpw > 0;
    const int vbase = warp * p.cv_cpw;
    const int vcnt = cv ? max(0, min(p.cv_cpw, p.n - vbase)) : 0;

    auto vorig_at = [&](int j) -> int { return p.perm_v ? static_cast<int>(lds_u16(vorig_a + 2 * j)) : j; };

    if (threadIdx.x == 0) tma_load_tables(smem + p.off_tables, p.tables, static_cast<uint32_t>(p.tables_bytes), mbar);

    // This CTA's queue: 32-syndrome chunks c, c+G, c+2G, ... of the batch (B < 2^31 - 32*G, host-checked).
    const int G = DUAL ? 2 * gridDim.x : gridDim.x, c = DUAL ? 2 * blockIdx.x + team : blockIdx.x;
    // (with a work list -- the syndromes the first-iteration filter left over -- queue entries index the list)
    const int Bn = p.list ? __ldg(p.list_count) : static_cast<int>(p.B);
    const int nchunks = (Bn + 31) >> 5;
    const int Q = (c < nchunks) ? (((nchunks - c + G - 1) / G) << 5) : 0;
    auto sid_of = [&](int q) -> int {
        const int sid = (((q >> 5) * G + c) << 5) + (q & 31);
        if (!(q < Q && sid < Bn)) return -1;
        return p.list ? __ldg(p.list + sid) : sid;
    };
    // wS moves staged syndromes into the lanes' own rows, wP prefetches the next queue window,
    // wO writes converged flags / iteration counts and keeps the counters.
    const int wS = 0, wP = (W > 1) ? 1 : 0, wO = (W > 2) ? 2 : 0;
    // Dynamic queue (p.queue_ctr != null): instead of the static share c, c+G, ... the prefetching warp claims the next
    // 32-syndrome chunk of the batch from a global counter whenever its sliding window reaches a new one, so CTAs whose
    // syndromes happened to be easy keep taking work and all CTAs end together.  dq_a / dq_b: the claimed chunks that the
    // window's two local chunks dq_k, dq_k + 1 stand for (warp wP only; the other warps read the staged indices).
    int dq_k = -1, dq_a = 0, dq_b = 0;
    auto claim = [&]() -> int {
        int g = 0;
        if (lane == 0) g = static_cast<int>(atomicAdd(p.queue_ctr, 1u));
        return __shfl_sync(0xffffffffu, g, 0);
    };
    auto prefetch = [&](int q_head, int buf) {            // stage[buf][w][r] <- syndrome words of entry q_head + r
        int sid;
        if (p.queue_ctr != nullptr) {
            const int k0 = q_head >> 5;
            if (dq_k < 0) { dq_a = claim(); dq_b = claim(); dq_k = 0; }
            while (dq_k < k0) { dq_a = dq_b; dq_b = claim(); ++dq_k; }
            const int q = q_head + lane;
            const int g = (q >> 5) == k0 ? dq_a : dq_b;
            const int idx = (g << 5) + (q & 31);
            sid = (g < nchunks && idx < Bn) ? (p.list ? __ldg(p.list + idx) : idx) : -1;
        } else {
            sid = sid_of(q_head + lane);
        }
        // (a refill reads the index from here instead of repeating the dependent work-list load in every warp)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(sidq_a + buf * 128 + lane * 4), "r"(sid) : "memory");
        if (sid >= 0) {
            const uint32_t *src = p.syn_words + static_cast<size_t>(sid) * p.SW;
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + goff + p.off_stage) + buf * p.SW * 32 + lane;
            for (int w = 0; w < p.SW; ++w) cp_async4(dst + w * 32, src + w);
        }
    };

    int q_head = 0, sbuf = 0;
    int sid = -1, iter = 0;
    bool active = false, fresh = false;
    uint32_t syn_a = syn_own;                             // where this lane's syndrome is read from in the check pass
    uint32_t res_a = res0;                                // this lane's live residual buffer
    ebits_t ebits = 0;                                    // decisions of the variables this warp owns (bit i <-> j = warp + i*W)
    unsigned long long n_done = 0, n_conv = 0, n_iters = 0;   // warp wO only

    if (DUAL && threadIdx.x < 4) team_done[threadIdx.x] = 0;   // team_done[2], team_seen[2]
    if (warp == wP) { prefetch(0, 0); cp_async_wait_all(); }
    __syncthreads();
    mbar_wait(mbar, 0);                                   // tables have landed
    if (DUAL && team == 1) token_pass();                  // team 0 takes the first turn

    // Lanes in `mask` take the next queue entries (identically in every warp: the lane state is replicated).
    auto refill = [&](uint32_t mask) {
        if ((mask >> lane) & 1u) {
            const int rank = __popc(mask & lt_mask);
            sid = static_cast<int>(lds_u32(sidq_a + sbuf * 128 + rank * 4));
            active = sid >= 0;
            fresh = active;
            iter = 0;
            ebits = 0;                                    // err .= 0 (reset!, :89)
            res_a = res_sum - res_a;                      // the other residual buffer
            if (active) {
                syn_a = stage_a + sbuf * (p.SW * 128) + rank * 4;
                if (warp == wS)
                    for (int w = 0; w < p.SW; ++w) {
                        const uint32_t v = lds_u32(syn_a + w * 128);
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(syn_own + w * 128), "r"(v) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(res_a + w * 128), "r"(v) : "memory");
                    }
            }
        }
        q_head += __popc(mask);
        sbuf ^= 1;
        if (warp == wP) prefetch(q_head, sbuf);           // lands before the barrier that precedes the next refill
    };

    refill(0xffffffffu);
    uint32_t any_active = __ballot_sync(0xffffffffu, active);

    const double p0 = p.p0;
    const double caux = p.check_aux;
    const bool regular_p0 = p.regular_p0;
    // optional phase timing (option "kernel_profile"): SM cycles per warp summed over all warps
    constexpr bool prof = PROF;                            // compiled into one extra instantiation only (bp_smem_inst.cuh)
    uint32_t pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;     // 32-bit sums: a warp's share of one launch stays far below 2^32 cycles
    auto tick = [&](int k) {
        if constexpr (prof) { const uint32_t t = static_cast<uint32_t>(clock64()); pc[k] += t - pt; pt = t; }
    };
    if constexpr (prof) pt = static_cast<uint32_t>(clock64());
    while (any_active != 0u) {
        // ------------------------------------------------------------------ check pass (:135-150)
        token_wait();
        if (active) {
            int i = warp;
            for (int g = 0; g < p.seg.ncseg; ++g) {
                const int deg = p.seg.cdeg[g], first = p.seg.cfirst[g], end = p.seg.cend[g], sb = p.seg.cslot[g];
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        uint32_t a = ml + (sb + (i - first) * D) * 256;                                          \
        if (p.perm_c) {                                                                          \
            for (; i < end; i += W, a += W * (D * 256)) {                                        \
                const int io = static_cast<int>(lds_u16(corig_a + 2 * i));                       \
                check_node<D>(a, (lds_u32(syn_a + (io >> 5) * 128) >> (io & 31)) & 1u, fresh, p0, caux); \
            }                                                                                    \
        } else {                                                                                 \
            for (; i < end; i += W, a += W * (D * 256))                                          \
                check_node<D>(a, (lds_u32(syn_a + (i >> 5) * 128) >> (i & 31)) & 1u, fresh, p0, caux); \
        }                                                                                        \
    }
                BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
                if (deg == 0 && i < end) i += ((end - i + W - 1) / W) * W;   // isolated checks: nothing to send
            }
        }
        fresh = false;
        syn_a = syn_own;                                   // (wS copied the staged syndrome into the lane's own rows at refill time)
        token_pass();
        tick(0);
        team_sync();
        tick(1);
        if (DUAL && warp == 0 && lane == 0) team_seen[team] = team_done[team ^ 1];   // read by the whole team after the next barrier
        // --------------------------------------------------------------- variable pass (:152-178)
        if (active) {
            ebits_t newbits = 0;
            const bool wr = p.ratio != nullptr && (!p.ratio_last_only || iter + 1 >= p.max_iters);
            if (cv) {
                // contiguous ownership: this warp's variables are vbase .. vbase + vcnt - 1 (bit i <-> vbase + i); four per trip
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        uint32_t ta = ve_a + warp * p.cv_stride;                                                 \
        int i = 0;                                                                               \
        if constexpr (D <= 4) {                                                                  \
            for (; i + 4 <= vcnt; i += 4, ta += 16 * D) {                                        \
                double R[4];                                                                     \
                var_group<D, 4>(ml, ta, p0, regular_p0, R);                                      \
                if (wr) {                                                                        \
                    double *rr = p.ratio + static_cast<size_t>(sid) * p.n + vbase + i;           \
                    rr[0] = R[0]; rr[1] = R[1]; rr[2] = R[2]; rr[3] = R[3];                      \
                }                                                                                \
                const uint32_t nib = (decide(R[0]) ? 1u : 0u) | (decide(R[1]) ? 2u : 0u) | (decide(R[2]) ? 4u : 0u) | (decide(R[3]) ? 8u : 0u); \
                newbits |= static_cast<ebits_t>(nib) << i;                                       \
            }                                                                                    \
        }                                                                                        \
        for (; i < vcnt; ++i, ta += 4 * D) {                                                     \
            double R[1];                                                                         \
            var_group<D, 1>(ml, ta, p0, regular_p0, R);                                          \
            if (wr) p.ratio[static_cast<size_t>(sid) * p.n + vbase + i] = R[0];                  \
            newbits |= static_cast<ebits_t>(decide(R[0]) ? 1u : 0u) << i;                       \
        }                                                                                        \
    }
                BP_DEGREE_SWITCH(p.uni_vdeg, BP_CASE, ;)
#undef BP_CASE
            } else {
            int j = warp, i = 0;
            for (int g = 0; g < p.seg.nvseg; ++g) {
                const int deg = p.seg.vdeg[g], first = p.seg.vfirst[g], end = p.seg.vend[g], eb = p.seg.vedge[g];
                if (deg == 0) {                                           // isolated variables: prior only
                    for (; j < end; j += W, ++i) {
                        if (wr) p.ratio[static_cast<size_t>(sid) * p.n + vorig_at(j)] = p0;
                        newbits |= static_cast<ebits_t>(decide(p0) ? 1u : 0u) << i;
                    }
                    continue;
                }
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        uint32_t vea = ve_a + 4 * (eb + (j - first) * D);                                        \
        const int step = W * D * 4;                                                              \
        if constexpr (D <= 4) {                                                                  \
            for (; j + W < end; j += 2 * W, i += 2, vea += 2 * step) {                           \
                uint32_t va[D], vb[D];                                                           \
                double ma[D], mb[D];                                                             \
                load_offsets<D>(va, vea);                                                        \
                load_offsets<D>(vb, vea + step);                                                 \
                _Pragma("unroll") for (int k = 0; k < D; ++k) ma[k] = ld_msg(ml + va[k]);        \
                _Pragma("unroll") for (int k = 0; k < D; ++k) mb[k] = ld_msg(ml + vb[k]);        \
                const double Ra = var_update<D>(ma, p0, regular_p0);                             \
                const double Rb = var_update<D>(mb, p0, regular_p0);                             \
                _Pragma("unroll") for (int k = 0; k < D; ++k) st_msg(ml + va[k], ma[k]);         \
                _Pragma("unroll") for (int k = 0; k < D; ++k) st_msg(ml + vb[k], mb[k]);         \
                if (wr) {                                                                        \
                    p.ratio[static_cast<size_t>(sid) * p.n + vorig_at(j)] = Ra;                  \
                    p.ratio[static_cast<size_t>(sid) * p.n + vorig_at(j + W)] = Rb;              \
                }                                                                                \
                newbits |= static_cast<ebits_t>((decide(Ra) ? 1u : 0u) | (decide(Rb) ? 2u : 0u)) << i; \
            }                                                                                    \
        }                                                                                        \
        for (; j < end; j += W, ++i, vea += step) {                                              \
            const double R = var_node<D>(ml, vea, p0, regular_p0);                               \
            if (wr) p.ratio[static_cast<size_t>(sid) * p.n + vorig_at(j)] = R;                   \
            newbits |= static_cast<ebits_t>(decide(R) ? 1u : 0u) << i;                          \
        }                                                                                        \
    }
                BP_DEGREE_SWITCH(deg, BP_CASE, ;)
#undef BP_CASE
            }
            }
            // Only variables whose decision flipped touch the residual syndrome s xor H*e (:180-181, kept
            // incrementally); every lane walks its own flips, no value comes back from the atomics.
            ebits_t f = ebits ^ newbits;
            ebits = newbits;
            tick(2);
            if (cv) {
#define BP_CASE(D)                                                                               \
    while (f) {                                  /* two flipped variables per trip: their table reads overlap */ \
        int b0, b1;                                                                              \
        if constexpr (EB64) b0 = __ffsll(static_cast<long long>(f)) - 1;                         \
        else b0 = __ffs(static_cast<int>(f)) - 1;                                                \
        f &= f - 1;                                                                              \
        const bool two = f != 0;                                                                 \
        if constexpr (EB64) b1 = two ? __ffsll(static_cast<long long>(f)) - 1 : b0;              \
        else b1 = two ? __ffs(static_cast<int>(f)) - 1 : b0;                                     \
        f &= f - 1;                              /* (0 & -1 stays 0) */                          \
        const uint32_t fa = vflip_a + 2 * D * (vbase + b0), fb = vflip_a + 2 * D * (vbase + b1); \
        uint32_t ea[D], eb[D];                                                                   \
        _Pragma("unroll") for (int k = 0; k < D; ++k) ea[k] = lds_u16(fa + 2 * k);               \
        _Pragma("unroll") for (int k = 0; k < D; ++k) eb[k] = lds_u16(fb + 2 * k);               \
        _Pragma("unroll") for (int k = 0; k < D; ++k)                                            \
            asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(res_a + (ea[k] & ~127u)), "r"(1u << (ea[k] & 31u)) : "memory"); \
        if (two) {                                                                               \
            _Pragma("unroll") for (int k = 0; k < D; ++k)                                        \
                asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(res_a + (eb[k] & ~127u)), "r"(1u << (eb[k] & 31u)) : "memory"); \
        }                                                                                        \
    }
                BP_DEGREE_SWITCH(p.uni_vdeg, BP_CASE, ;)
#undef BP_CASE
            }
            while (f) {
                int b;
                if constexpr (EB64) b = __ffsll(static_cast<long long>(f)) - 1;
                else b = __ffs(static_cast<int>(f)) - 1;
                f &= f - 1;
                const int jj = warp + b * W;
                int e0, e1;
                if (p.uni_vdeg) { e0 = jj * p.uni_vdeg; e1 = e0 + p.uni_vdeg; }
                else { e0 = static_cast<int>(lds_u16(colptr_a + 2 * jj)); e1 = static_cast<int>(lds_u16(colptr_a + 2 * jj + 2)); }
                for (int e = e0; e < e1; ++e) {
                    const uint32_t ent = lds_u16(vflip_a + 2 * e);                    // (check/32)*128 + check%32
                    const uint32_t addr = res_a + (ent & ~127u);
                    const uint32_t bit = 1u << (ent & 31u);
                    asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(addr), "r"(bit) : "memory");
                }
            }
        }
        tick(3);
        if (warp == wP) cp_async_wait_all();                  // staged window complete before anyone reads it
        team_sync();
        tick(4);
        if constexpr (DUAL) other_done = team_seen[team] != 0;
        // ---------------------------------------- syndrome re-check, early stop, refill (:180-184)
        uint32_t r = 0;
        for (int w = 0; w < p.SW; ++w) r |= lds_u32(res_a + w * 128);
        if (active) ++iter;
        const bool conv = active && r == 0u;
        const bool done = active && ((p.early_stop && conv) || iter >= p.max_iters);
        const uint32_t done_mask = __ballot_sync(0xffffffffu, done);
        if (done) {
            // errors[:, sid] = guess (:227): every warp ORs the set bits it owns into the pre-zeroed packed row
            // p.out_bits: the output is the caller's bit stream itself (Julia BitMatrix: bit sid*n + j), so that no
            // conversion kernel stands between this kernel and the copy to the host; else packed rows of NW words
            const unsigned long long obase = p.out_bits ? static_cast<unsigned long long>(sid) * static_cast<unsigned long long>(p.n)
                                                        : static_cast<unsigned long long>(sid) * static_cast<unsigned long long>(p.NW) * 32ull;
            ebits_t b = ebits;
            if (cv) {                                         // the field is bits vbase .. vbase+vcnt-1 of the row
                const unsigned long long o = obase + static_cast<unsigned long long>(vbase);
                uint32_t *w0 = p.err_words + (o >> 5);
                const int lo = static_cast<int>(o & 31ull);
                const unsigned long long v = static_cast<unsigned long long>(b) << lo;
                const uint32_t v0 = static_cast<uint32_t>(v), v1 = static_cast<uint32_t>(v >> 32);
                if (v0) atomicOr(w0, v0);
                if (v1) atomicOr(w0 + 1, v1);
                if constexpr (EB64) {
                    const uint32_t v2 = lo ? static_cast<uint32_t>(static_cast<unsigned long long>(b) >> (64 - lo)) : 0u;
                    if (v2) atomicOr(w0 + 2, v2);
                }
                b = 0;
            }
            while (b) {
                int bi;
                if constexpr (EB64) bi = __ffsll(static_cast<long long>(b)) - 1;
                else bi = __ffs(static_cast<int>(b)) - 1;
                b &= b - 1;
                const unsigned long long o = obase + static_cast<unsigned long long>(vorig_at(warp + bi * W));
                atomicOr(p.err_words + (o >> 5), 1u << static_cast<int>(o & 31ull));
            }
            if (warp == wO) {
                p.conv[sid] = conv ? 1 : 0;
                if (p.iters) p.iters[sid] = iter;
                n_done += 1; n_conv += conv ? 1 : 0; n_iters += iter;
            }
        }
        tick(5);
        if (done_mask) {
            refill(done_mask);
            any_active = __ballot_sync(0xffffffffu, active);
        }
        tick(6);
        if constexpr (prof) pc[7] += 1;
    }
    if constexpr (DUAL) {                                  // this team is finished: one last, empty turn
        token_wait();
        if (warp == 0 && lane == 0) { team_done[team] = 1; __threadfence_block(); }   // raised before this thread's arrival below
        token_pass();
    }
    if (prof && p.prof != nullptr && lane == 0)
        for (int k = 0; k < 8; ++k) atomicAdd(p.prof + k, static_cast<unsigned long long>(pc[k]));

    if (warp == wO && p.counters) {
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

}  // inline namespace BP_VNS
}  // namespace bp
