// bp_kernel.cuh -- the persistent BP kernel: one launch runs every iteration of every syndrome.
//
// Replaces the whole of decode!/batchdecode! (/root/reference/src/decoders/belief_propagation.jl
// :121-188, :220-231).
//
// Mapping
//   CTA  = 32 "lanes" (syndrome slots) x W warps.  Lane l of EVERY warp works on the syndrome
//          currently held in slot l; warp w owns checks w, w+W, ... in the check pass and
//          variables w, w+W, ... in the variable pass.  All node indices are therefore
//          warp-uniform (broadcast table reads, no divergence) and every message access is
//          msg[slot_edge][lane]: 256 contiguous bytes per warp -- bank-conflict free in shared
//          memory, fully coalesced in HBM.
//   Messages: ONE in-place array per CTA, check-major edge order (a check's edges are
//          contiguous); the check pass turns bit->check ratios into check->bit ratios in place,
//          the variable pass turns them back.
//   Where things live (template MODE):
//          0  family SMEM  : messages, syndrome state and edge tables in shared memory
//                            (surface d<=15, [[144,12,12]] gross code, ...; FP64-pipe bound)
//          1  family GLOBAL: messages in HBM/L2 (E*256 B per CTA), state and tables in shared
//                            memory (HGP-1600, Gallager n=1000, ...)
//          2  family GLOBAL: messages, syndrome state and (32-bit) tables in HBM/L2
//                            (Gallager n=100k: 77 MB of messages per CTA, HBM-bandwidth bound)
//   Hard decisions and the syndrome re-check (belief_propagation.jl:164-168,180-184) are kept
//          incrementally: each warp holds the decisions of its variables as a bit field in a
//          register, resid = s xor H*e bit-packed, nnz = popcount(resid).  A variable whose
//          decision flips XORs the bits of its checks in resid (atomics, lanes walk their own
//          flips).  converged <=> nnz == 0.  A finished lane ORs its set bits into the
//          pre-zeroed packed output row.
//   Early termination / compaction: every lane has its own iteration counter.  A lane whose
//          syndrome converged (or hit max_iters) writes its outputs and immediately takes the
//          next syndrome of the CTA's queue, so no lane idles on finished syndromes.  Fresh
//          lanes read the prior p/(1-p) instead of stored messages (initialisation :127-131
//          without a store pass).
//   Shared-memory edge tables are fetched with one TMA bulk copy (cp.async.bulk + mbarrier); the
//          next 32 queued syndromes are prefetched with cp.async (modes 0/1).
#pragma once
#include <type_traits>

#include "bp_math.cuh"

// mode 2 software pipelining of table / state loads from global memory (experiments: -DBP_M2_VAR_PIPE=0 etc.)
#ifndef BP_M2_VAR_PIPE
#define BP_M2_VAR_PIPE 0
#endif
#ifndef BP_M2_SYN_PRE
#define BP_M2_SYN_PRE 0
#endif

namespace bp {

// Node order inside the kernels is degree-sorted (host: build_graph): nodes of equal degree form
// contiguous segments [first, end) so the degree dispatch happens once per segment, and slot /
// table addresses inside a segment are affine in the node index.
constexpr int kMaxSeg = 8;
struct Segments {
    int ncseg, nvseg;                                 // 0 = too many distinct degrees: per-node path
    int cdeg[kMaxSeg], cfirst[kMaxSeg], cend[kMaxSeg], cslot[kMaxSeg];   // cslot = first message slot of the segment
    int vdeg[kMaxSeg], vfirst[kMaxSeg], vend[kMaxSeg], vedge[kMaxSeg];   // vedge = first edge-table entry of the segment
};

struct KernelParams {
    int s, n, E;
    int uni_cdeg, uni_vdeg;   // common check / variable degree if the code is regular in it (1..12), else 0
    int SW, NW;               // uint32 words per packed syndrome / error row
    int max_iters;
    int early_stop;           // 1 = reference semantics
    double p0;                // exact: per / (1 - per); min-sum: log((1 - per) / per)
    double check_aux;         // min-sum: normalisation factor applied to the check->variable magnitude
    int regular_p0;           // p0 is a positive normal double (no NaN clamp can fire on finite messages)
    long long B;
    const uint32_t *syn_words;    // [B][SW]
    uint32_t *err_words;          // [B][NW], zero on entry
    uint8_t *conv;                // [B]
    int32_t *iters;               // [B] or null
    double *ratio;                // [B][n] or null
    int ratio_last_only;          // 1: write ratios only in iteration max_iters
    unsigned long long *counters; // [4] or null
    const int *list;              // bp_smem_kernel: null, or the syndromes to decode (indices into the batch) ...
    const int *list_count;        // ... and how many (device memory; written by first_iter_filter_kernel)
    unsigned long long *prof;     // bp_smem_kernel: null, or [8] phase cycle sums (check, wait B1, variable, flips, wait B2, done+emit, refill, iterations)
    // narrow tables (modes 0/1), copied to shared memory:
    const unsigned char *tables;  // global blob: rowptr u16[s+1] | colptr u16[n+1] | ve_off u32[E] | vflip u16[E]
    int tables_bytes;             // multiple of 16
    int off_colptr, off_ve, off_vflip;   // byte offsets inside the blob (rowptr at 0)
    int off_corig, off_vorig;            // u16 original ids of the kernels' check / variable order (if permuted)
    int group_stride;                    // bp_smem_kernel<DUAL>: byte distance between the two teams' groups of arrays (0: one group)
    int cv_cpw, cv_stride;               // bp_smem_kernel, contiguous variable ownership: variables per warp (0 = interleaved), bytes per warp in the ve table
    int perm_c, perm_v;                  // node order differs from the caller's (degree-sorted)
    Segments seg;
    // wide tables (mode 2), read from global memory:
    const int *g_corig, *g_vorig;        // [s], [n] original ids
    const int *g_rowptr, *g_colptr;      // [s+1], [n+1]
    const uint32_t *g_ve_off;            // [E] slot * 256
    const uint32_t *g_vflip;             // [E] (check/32)*128 + check%32
    // global stores (modes 1/2): per CTA E*32 doubles of messages; mode 2 also 2*SW*32 words of state
    double *msg_global;
    uint32_t *state_global;
    // decision bit fields when a warp owns more than 64 variables: nfw words per thread, laid out
    // [word][thread]; in shared memory (off_efield) or, if efield_global != null, per CTA in HBM
    int nfw;
    uint32_t *efield_global;
    // modes 1/2: per-warp shared-memory ring that message rows are prefetched into with cp.async
    // (pd + 1 slots of ring_slot_bytes each; pd = prefetch distance in nodes, 0 = no staging)
    int pd, ring_slot_bytes, ring_warp_bytes;
    // shared-memory carve-up (byte offsets from the dynamic smem base)
    int off_syn, off_resid, off_stage, off_nnz, off_tables, off_mbar, off_efield, off_ring;
    int off_sidq;             // bp_smem_kernel: [2][32] syndrome indices of the staged queue window
    unsigned int *queue_ctr;  // bp_smem_kernel: null (static shares of the batch per CTA) or a zeroed counter of claimed 32-syndrome chunks
    int out_bits;             // bp_smem_kernel: err_words is a bit stream (bit sid*n + j) instead of rows of NW words
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most n (0..kMaxPrefetch) of the most recent groups are still in flight
constexpr int kMaxPrefetch = 6;
__device__ __forceinline__ void cp_async_wait_pending(int n)
{
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    }
}


// ---- message / table accessors.  Shared-window handles are 32-bit shared-space addresses
// computed once (no generic->shared conversions in the hot loops); global handles are pointers.
template <int OFF = 0>
__device__ __forceinline__ double ld_msg(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF = 0>
__device__ __forceinline__ void st_msg(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
template <int OFF = 0>
__device__ __forceinline__ double ld_msg(unsigned char *p) { return *reinterpret_cast<double *>(p + OFF); }
template <int OFF = 0>
__device__ __forceinline__ void st_msg(unsigned char *p, double v) { *reinterpret_cast<double *>(p + OFF) = v; }

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
// k-th slot offset of a variable: shared table (address) or global table (pointer)
template <int K>
__device__ __forceinline__ uint32_t ld_off(uint32_t vea) { return lds_u32<K * 4>(vea); }
template <int K>
__device__ __forceinline__ uint32_t ld_off(const uint32_t *vep) { return __ldg(vep + K); }

template <int D, class MH, int K = 0>
__device__ __forceinline__ void load_row(double (&m)[D], MH a)
{
    if constexpr (K < D) {
        m[K] = ld_msg<K * 256>(a);
        load_row<D, MH, K + 1>(m, a);
    }
}
template <int D, class MH, int K = 0>
__device__ __forceinline__ void store_row(const double (&m)[D], MH a)
{
    if constexpr (K < D) {
        st_msg<K * 256>(a, m[K]);
        store_row<D, MH, K + 1>(m, a);
    }
}
template <int D, class TH, int K = 0>
__device__ __forceinline__ void load_offsets(uint32_t (&v)[D], TH a)
{
    if constexpr (K < D) {
        v[K] = ld_off<K>(a);
        load_offsets<D, TH, K + 1>(v, a);
    }
}

inline namespace BP_VNS {

// One check node of degree D whose D message slots start at `a` (this lane's column,
// consecutive slots 256 B apart).
template <int D, class MH>
__device__ __forceinline__ void check_node(MH a, bool neg, bool fresh, double p0, double aux)
{
    double m[D];
    load_row<D>(m, a);
    if (fresh) {
#pragma unroll
        for (int k = 0; k < D; ++k) m[k] = p0;
    }
    check_update<D>(m, neg, aux);
    store_row<D>(m, a);
}

// One variable node of degree D; `vea` = handle of its D slot offsets, `ml` = this lane's
// message column.  Returns the posterior ratio.
template <int D, class MH, class TH>
__device__ __forceinline__ double var_node(MH ml, TH vea, double p0, bool regular_p0)
{
    uint32_t v[D];
    double m[D];
    load_offsets<D>(v, vea);
#pragma unroll
    for (int k = 0; k < D; ++k) m[k] = ld_msg(ml + v[k]);
    const double R = var_update<D>(m, p0, regular_p0);
#pragma unroll
    for (int k = 0; k < D; ++k) st_msg(ml + v[k], m[k]);
    return R;
}

// Staged forms (modes 1/2): the node's rows were prefetched into a shared-memory ring slot
// (`ra`, this lane's column); results go straight back to global memory.
template <int D>
__device__ __forceinline__ void check_node_staged(uint32_t ra, unsigned char *ga, bool neg, bool fresh, double p0, double aux)
{
    double m[D];
    load_row<D>(m, ra);
    if (fresh) {
#pragma unroll
        for (int k = 0; k < D; ++k) m[k] = p0;
    }
    check_update<D>(m, neg, aux);
    store_row<D>(m, ga);
}
template <int D, class TH>
__device__ __forceinline__ double var_node_staged(uint32_t ra, unsigned char *ml, TH vea, double p0, bool regular_p0)
{
    uint32_t v[D];
    double m[D];
    load_offsets<D>(v, vea);
    load_row<D>(m, ra);
    const double R = var_update<D>(m, p0, regular_p0);
#pragma unroll
    for (int k = 0; k < D; ++k) st_msg(ml + v[k], m[k]);
    return R;
}

// ... with the D slot offsets already in registers (mode 2 loads them one node ahead: they come from global memory)
template <int D>
__device__ __forceinline__ double var_node_staged_off(uint32_t ra, unsigned char *ml, const uint32_t (&v)[D], double p0, bool regular_p0)
{
    double m[D];
    load_row<D>(m, ra);
    const double R = var_update<D>(m, p0, regular_p0);
#pragma unroll
    for (int k = 0; k < D; ++k) st_msg(ml + v[k], m[k]);
    return R;
}

// Residual-syndrome update for the flipped variables of a lane when every variable has degree D:
// variable j's D (word, bit) entries sit at vflip[j*D ..], no column-pointer reads.  `resid_lane`
// points at this lane's column of the residual words ([word][32] layout).  Returns the change of
// the number of unsatisfied checks.
template <int D, class VF>
__device__ __forceinline__ int flip_walk_uniform(uint32_t f, int ibase, int warp, int W, VF vflip_at, uint32_t *resid_lane)
{
    int delta = 0;
    while (f) {
        const int b = __ffs(static_cast<int>(f)) - 1;
        f &= f - 1;
        const int e0 = (warp + (ibase + b) * W) * D;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint32_t ent = vflip_at(e0 + k);                        // (check/32)*128 + check%32
            const uint32_t old = atomicXor(resid_lane + (ent >> 7) * 32, 1u << (ent & 31u));
            delta += 1 - 2 * static_cast<int>((old >> (ent & 31u)) & 1u);
        }
    }
    return delta;
}

template <int MODE, bool BIG, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bp_persistent_kernel(const __grid_constant__ KernelParams p)
{
    constexpr bool kMsgShared = MODE == 0;
    constexpr bool kStateShared = MODE <= 1;
    using MH = typename std::conditional<kMsgShared, uint32_t, unsigned char *>::type;   // message column handle
    using TH = typename std::conditional<kStateShared, uint32_t, const uint32_t *>::type;  // slot-offset table handle

    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t sbase = smem_u32(smem);

    // ---- where this CTA's data lives
    MH ml;                                            // this lane's message column
    unsigned char *msg_generic;                       // same, as a generic pointer (local-memory degree path)
    if constexpr (kMsgShared) {
        ml = sbase + lane * 8;
        msg_generic = smem + lane * 8;
    } else {
        msg_generic = reinterpret_cast<unsigned char *>(p.msg_global + static_cast<size_t>(blockIdx.x) * p.E * 32) + lane * 8;
        ml = msg_generic;
    }
    uint32_t *syn, *resid;                            // [SW][32] each
    if constexpr (kStateShared) {
        syn = reinterpret_cast<uint32_t *>(smem + p.off_syn);
        resid = reinterpret_cast<uint32_t *>(smem + p.off_resid);
    } else {
        syn = p.state_global + static_cast<size_t>(blockIdx.x) * 2 * p.SW * 32;
        resid = syn + static_cast<size_t>(p.SW) * 32;
    }
    // decision fields in memory (only when a warp owns more than 64 variables)
    const bool use_regs = p.n <= 64 * W;
    uint32_t *efield = p.efield_global ? p.efield_global + static_cast<size_t>(blockIdx.x) * p.nfw * blockDim.x + threadIdx.x
                                       : reinterpret_cast<uint32_t *>(smem + p.off_efield) + threadIdx.x;
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem + p.off_stage);
    int *nnz = reinterpret_cast<int *>(smem + p.off_nnz);          // [2][32]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + p.off_mbar);
    const uint32_t syn_a = sbase + p.off_syn + lane * 4;
    const uint32_t rowptr_a = sbase + p.off_tables;
    const uint32_t colptr_a = sbase + p.off_tables + p.off_colptr;
    const uint32_t ve_a = sbase + p.off_tables + p.off_ve;         // u32 byte offset of each edge's slot row
    const uint32_t vflip_a = sbase + p.off_tables + p.off_vflip;   // u16 residual word offset | bit of each edge's check

    auto rowptr_at = [&](int i) -> int {
        if constexpr (kStateShared) return static_cast<int>(lds_u16(rowptr_a + 2 * i));
        else return __ldg(p.g_rowptr + i);
    };
    auto colptr_at = [&](int j) -> int {
        if constexpr (kStateShared) return static_cast<int>(lds_u16(colptr_a + 2 * j));
        else return __ldg(p.g_colptr + j);
    };
    auto vflip_at = [&](int e) -> uint32_t {
        if constexpr (kStateShared) return lds_u16(vflip_a + 2 * e);
        else return __ldg(p.g_vflip + e);
    };
    auto ve_handle = [&](int e) -> TH {               // handle of the slot offsets starting at edge e
        if constexpr (kStateShared) return ve_a + 4 * e;
        else return p.g_ve_off + e;
    };
    const uint32_t corig_a = sbase + p.off_tables + p.off_corig;
    const uint32_t vorig_a = sbase + p.off_tables + p.off_vorig;
    auto corig_at = [&](int i) -> int {               // original index of the kernels' i-th check
        if (!p.perm_c) return i;
        if constexpr (kStateShared) return static_cast<int>(lds_u16(corig_a + 2 * i));
        else return __ldg(p.g_corig + i);
    };
    auto vorig_at = [&](int j) -> int {
        if (!p.perm_v) return j;
        if constexpr (kStateShared) return static_cast<int>(lds_u16(vorig_a + 2 * j));
        else return __ldg(p.g_vorig + j);
    };
    auto syn_bit_direct = [&](int i) -> bool {        // syndrome bit of original check i
        if constexpr (kStateShared) return (lds_u32(syn_a + (i >> 5) * 128) >> (i & 31)) & 1u;
        else return (syn[(i >> 5) * 32 + lane] >> (i & 31)) & 1u;
    };
    auto syn_bit = [&](int ik) -> bool {              // syndrome bit of the kernels' ik-th check
        const int i = corig_at(ik);
        if constexpr (kStateShared) return (lds_u32(syn_a + (i >> 5) * 128) >> (i & 31)) & 1u;
        else return (syn[(i >> 5) * 32 + lane] >> (i & 31)) & 1u;
    };

    if constexpr (kStateShared) {
        if (threadIdx.x == 0)
            tma_load_tables(smem + p.off_tables, p.tables, static_cast<uint32_t>(p.tables_bytes), mbar);
    }

    // This CTA's queue: 32-syndrome chunks c, c+G, c+2G, ... of the batch.
    const long long G = gridDim.x, c = blockIdx.x;
    const long long nchunks = (p.B + 31) >> 5;
    const long long my_chunks = (c < nchunks) ? (nchunks - c + G - 1) / G : 0;
    const long long Q = my_chunks << 5;
    // Dynamic queue (p.queue_ctr != null): warp wP claims the next chunk of the batch from a global counter whenever the
    // sliding window reaches a new one and leaves its id in dq[k & 3] (k = local chunk number); every warp translates queue
    // positions through that table.  Four slots: the refill of window [k0, k0+1] may still be reading while wP already
    // writes the ids of [k0+1, k0+2] for the next one.
    volatile int *dq = reinterpret_cast<volatile int *>(smem + p.off_mbar + 8);
    const bool dynq = p.queue_ctr != nullptr;
    int dq_hi = -1;                                   // highest local chunk claimed so far (warp wP only)
    auto sid_of = [&](long long q) -> long long {
        if (dynq) {
            const long long g = dq[(q >> 5) & 3];
            const long long sid = (g << 5) + (q & 31);
            return (g < nchunks && sid < p.B) ? sid : -1;
        }
        const long long sid = (((q >> 5) * G + c) << 5) + (q & 31);
        return (q < Q && sid < p.B) ? sid : -1;
    };
    // Bookkeeping of the done-detection phase is spread over three warps so that no single warp
    // holds the others up at the barrier: wS moves staged syndromes into slots, wP prefetches the
    // next queue window (double-buffered stage, so it may overwrite while wS still reads), wO
    // writes converged flags / iteration counts and keeps the counters.
    const int wS = 0, wP = (W > 1) ? 1 : 0, wO = (W > 2) ? 2 : 0;
    int sbuf = 0;                                 // stage buffer the next refill reads
    auto prefetch = [&](long long q_head, int buf) {      // stage[buf][w][r] <- syndrome words of entry q_head + r
        if (dynq) {                                           // (warp wP) chunks of the window [q_head, q_head + 32)
            const int k1 = static_cast<int>(q_head >> 5) + 1;
            while (dq_hi < k1) {
                ++dq_hi;
                if (lane == 0) dq[dq_hi & 3] = static_cast<int>(atomicAdd(p.queue_ctr, 1u));
            }
            __syncwarp();
        }
        if constexpr (kStateShared) {
            const long long sid = sid_of(q_head + lane);
            if (sid >= 0)
                for (int w = 0; w < p.SW; ++w)
                    cp_async4(&stage[(buf * p.SW + w) * 32 + lane], p.syn_words + sid * p.SW + w);
        }
    };

    long long q_head = 0;
    long long sid = -1;
    int iter = 0;
    bool active = false, fresh = false;
    int par = 0;                                  // which nnz buffer the coming iteration updates
    // Hard decisions of the variables this warp owns (j = warp + i*W  <->  bit i), one bit set per
    // variable currently decided 1 (codes with at most 64 variables per warp; else `efield`).
    unsigned long long ebits = 0;
    unsigned long long n_done = 0, n_conv = 0, n_iters = 0;   // warp wO only

    if (warp == wP) { prefetch(0, sbuf); cp_async_wait_all(); }
    __syncthreads();
    if constexpr (kStateShared) mbar_wait(mbar, 0);           // tables have landed

    // Lanes in `mask` take the next queue entries.  Executed identically by every warp
    // (register state is replicated); warp wS additionally moves the syndrome in.  The staged
    // window was completed by wP before the preceding barrier.
    auto refill = [&](uint32_t mask, int nnz_buf) {
        if ((mask >> lane) & 1u) {
            const int rank = __popc(mask & lt_mask);
            sid = sid_of(q_head + rank);
            active = sid >= 0;
            fresh = active;
            iter = 0;
            ebits = 0;                                                      // err .= 0 (reset!, :89)
            if (!use_regs)
                for (int k = 0; k < p.nfw; ++k) efield[static_cast<size_t>(k) * blockDim.x] = 0u;
            if (warp == wS && active) {
                int cnt = 0;
                for (int w = 0; w < p.SW; ++w) {
                    uint32_t v;
                    if constexpr (kStateShared) v = stage[(sbuf * p.SW + w) * 32 + rank];
                    else v = p.syn_words[sid * p.SW + w];
                    syn[w * 32 + lane] = v;
                    resid[w * 32 + lane] = v;
                    cnt += __popc(v);
                }
                nnz[nnz_buf * 32 + lane] = cnt;
            }
        }
        q_head += __popc(mask);
        sbuf ^= 1;
        if (warp == wP) prefetch(q_head, sbuf);               // lands before the next refill's barrier
    };

    refill(0xffffffffu, par);
    __syncthreads();

    const double p0 = p.p0;
    const double caux = p.check_aux;                  // min-sum: normalisation factor; exact variant: unused
    const bool regular_p0 = p.regular_p0;
    while (__ballot_sync(0xffffffffu, active) != 0u) {
        // ------------------------------------------------------------------ check pass (:135-150)
        // warp w owns checks w, w+W, ...; syndrome bit of check i = bit i%32 of syn[i/32][lane]
        bool staged = false;
        if constexpr (!kMsgShared) staged = p.pd > 0;
        if (staged) {
            // Messages live in HBM/L2: the whole warp (active lanes or not) streams the rows of its
            // next pd checks into a shared-memory ring with 16-byte cp.async.cg copies (two 256 B rows
            // per instruction), so pd nodes' worth of loads are always in flight per warp.
            if constexpr (!kMsgShared) {
              if (p.seg.ncseg > 0) {
                // degree segments: compile-time degree, affine row addresses, the copy loop unrolled
                const int nslot = p.pd + 1;
                const uint32_t ring_c = sbase + p.off_ring + warp * p.ring_warp_bytes + (lane & 15) * 16;   // copy view
                const uint32_t ring_l = sbase + p.off_ring + warp * p.ring_warp_bytes + lane * 8;           // lane view
                unsigned char *cta_msg = msg_generic - lane * 8;
                int i = warp;
                for (int g = 0; g < p.seg.ncseg; ++g) {
                    const int deg = p.seg.cdeg[g], first = p.seg.cfirst[g], end = p.seg.cend[g], sb = p.seg.cslot[g];
                    if (deg == 0) { if (i < end) i += ((end - i + W - 1) / W) * W; continue; }
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        const unsigned char *isrc = cta_msg + static_cast<size_t>(sb + (i - first) * D) * 256 + (lane & 15) * 16; \
        unsigned char *ga = msg_generic + static_cast<size_t>(sb + (i - first) * D) * 256;       \
        int ii = i, islot = 0, cslot = 0;                                                        \
        auto issue = [&]() {                                                                     \
            if (ii < end) {                                                                      \
                _Pragma("unroll") for (int k = 0; k < D; k += 2)                                 \
                    if (k + (lane >> 4) < D)                                                     \
                        cp_async16(ring_c + islot * p.ring_slot_bytes + (k + (lane >> 4)) * 256, isrc + (k + (lane >> 4)) * 256); \
            }                                                                                    \
            cp_async_commit();                                                                   \
            ii += W; isrc += static_cast<size_t>(W) * (D * 256);                                 \
            islot = (islot + 1 == nslot) ? 0 : islot + 1;                                        \
        };                                                                                       \
        /* two checks per trip (computed one after the other: the trip's bookkeeping is shared, not the registers) when \
           the ring slot holds the rows of both */                                               \
        if constexpr (D <= 6) if (2 * D * 256 <= p.ring_slot_bytes) {                            \
            auto issue2 = [&]() {                                                                \
                if (ii + W < end) {                                                              \
                    _Pragma("unroll") for (int k = 0; k < 2 * D; k += 2) {                       \
                        const int r = k + (lane >> 4);                                           \
                        if (r < 2 * D)                                                           \
                            cp_async16(ring_c + islot * p.ring_slot_bytes + r * 256,             \
                                       isrc + (r < D ? r * 256 : static_cast<size_t>(W) * (D * 256) + (r - D) * 256)); \
                    }                                                                            \
                }                                                                                \
                cp_async_commit();                                                               \
                ii += 2 * W; isrc += static_cast<size_t>(2 * W) * (D * 256);                     \
                islot = (islot + 1 == nslot) ? 0 : islot + 1;                                    \
            };                                                                                   \
            for (int t = 0; t < p.pd; ++t) issue2();                                             \
            for (; i + W < end; i += 2 * W, ga += static_cast<size_t>(2 * W) * (D * 256)) {      \
                issue2();                                                                        \
                cp_async_wait_pending(p.pd);                                                     \
                __syncwarp();                                                                    \
                if (active) {                                                                    \
                    const uint32_t ra_ = ring_l + cslot * p.ring_slot_bytes;                     \
                    check_node_staged<D>(ra_, ga, syn_bit(i), fresh, p0, caux);                  \
                    check_node_staged<D>(ra_ + D * 256, ga + static_cast<size_t>(W) * (D * 256), syn_bit(i + W), fresh, p0, caux); \
                }                                                                                \
                __syncwarp();                                                                    \
                cslot = (cslot + 1 == nslot) ? 0 : cslot + 1;                                    \
            }                                                                                    \
            cp_async_wait_all();                                                                 \
            ii = i; islot = 0; cslot = 0;                                                        \
            isrc = cta_msg + static_cast<size_t>(sb + (i - first) * D) * 256 + (lane & 15) * 16; \
        }                                                                                        \
        for (int t = 0; t < p.pd; ++t) issue();                                                  \
        /* mode 2 keeps the syndrome in global memory (L2): the word of the NEXT check is fetched one trip ahead */ \
        uint32_t sw_cur = 0, sw_nxt = 0;                                                         \
        const bool sw_pre = BP_M2_SYN_PRE && !kStateShared && !p.perm_c;                                        \
        if constexpr (!kStateShared) { if (sw_pre && i < end) sw_cur = syn[(i >> 5) * 32 + lane]; } \
        for (; i < end; i += W, ga += static_cast<size_t>(W) * (D * 256)) {                      \
            issue();                                                                             \
            if constexpr (!kStateShared) { if (sw_pre && i + W < end) sw_nxt = syn[((i + W) >> 5) * 32 + lane]; } \
            cp_async_wait_pending(p.pd);                                                         \
            __syncwarp();                                                                        \
            if (active) check_node_staged<D>(ring_l + cslot * p.ring_slot_bytes, ga, sw_pre ? ((sw_cur >> (i & 31)) & 1u) != 0u : syn_bit(i), fresh, p0, caux); \
            __syncwarp();                                                                        \
            sw_cur = sw_nxt;                                                                     \
            cslot = (cslot + 1 == nslot) ? 0 : cslot + 1;                                        \
        }                                                                                        \
        cp_async_wait_all();                                                                     \
    }
                    BP_DEGREE_SWITCH(
                        deg, BP_CASE, if (BIG) {
                            for (; i < end; i += W) {
                                if (active) {
                                    double *base = reinterpret_cast<double *>(msg_generic + static_cast<size_t>(sb + (i - first) * deg) * 256);
                                    check_update_big([&](int k) -> double & { return base[k * 32]; }, deg, syn_bit(i), fresh, p0);
                                }
                            }
                        })
#undef BP_CASE
                }
              } else {
                const int nslot = p.pd + 1;
                const uint32_t ring_w = sbase + p.off_ring + warp * p.ring_warp_bytes;
                unsigned char *cta_msg = msg_generic - lane * 8;
                auto span = [&](int i, int &rp, int &deg) {
                    if (p.uni_cdeg) { deg = p.uni_cdeg; rp = i * deg; }
                    else { rp = rowptr_at(i); deg = rowptr_at(i + 1) - rp; }
                };
                auto issue = [&](int i, int slot) {
                    if (i < p.s) {
                        int rp, deg;
                        span(i, rp, deg);
                        if (deg <= kMaxRegDegree) {
                            const uint32_t dst = ring_w + slot * p.ring_slot_bytes + (lane & 15) * 16;
                            const unsigned char *src = cta_msg + static_cast<size_t>(rp) * 256 + (lane & 15) * 16;
                            for (int k = lane >> 4; k < deg; k += 2) cp_async16(dst + k * 256, src + static_cast<size_t>(k) * 256);
                        }
                    }
                    cp_async_commit();
                };
                int islot = 0;
                for (int t = 0; t < p.pd; ++t) { issue(warp + t * W, islot); islot = (islot + 1 == nslot) ? 0 : islot + 1; }
                int cslot = 0;
                for (int i = warp; i < p.s; i += W) {
                    issue(i + p.pd * W, islot);
                    islot = (islot + 1 == nslot) ? 0 : islot + 1;
                    cp_async_wait_pending(p.pd);
                    __syncwarp();
                    if (active) {
                        int rp, deg;
                        span(i, rp, deg);
                        const bool neg = syn_bit(i);
                        const uint32_t ra = ring_w + cslot * p.ring_slot_bytes + lane * 8;
                        unsigned char *ga = msg_generic + static_cast<size_t>(rp) * 256;
#define BP_CASE(D) check_node_staged<D>(ra, ga, neg, fresh, p0, caux)
                        BP_DEGREE_SWITCH(
                            deg, BP_CASE, if (BIG) {
                                double *base = reinterpret_cast<double *>(ga);
                                check_update_big([&](int k) -> double & { return base[k * 32]; }, deg, neg, fresh, p0);
                            })
#undef BP_CASE
                    }
                    __syncwarp();                       // slot is free for the copy issued next iteration
                    cslot = (cslot + 1 == nslot) ? 0 : cslot + 1;
                }
                cp_async_wait_all();
              }
            }
        } else if (active) {
            if (p.uni_cdeg) {
                // every check has the same degree (and the node order is the caller's): slots of
                // check i start at i*D, no table or segment reads
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        MH a = ml + warp * (D * 256);                                                            \
        for (int i = warp; i < p.s; i += W, a += W * (D * 256)) check_node<D>(a, syn_bit_direct(i), fresh, p0, caux); \
    }
                BP_DEGREE_SWITCH(p.uni_cdeg, BP_CASE, ;)
#undef BP_CASE
            } else if (p.seg.ncseg > 0) {
                // segments of equal degree: slots of the segment's u-th check start at cslot + u*D
                int i = warp;
                for (int g = 0; g < p.seg.ncseg; ++g) {
                    const int deg = p.seg.cdeg[g], first = p.seg.cfirst[g], end = p.seg.cend[g], sb = p.seg.cslot[g];
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        MH a = ml + (sb + (i - first) * D) * 256;                                                \
        if (p.perm_c)                                                                            \
            for (; i < end; i += W, a += W * (D * 256)) check_node<D>(a, syn_bit(i), fresh, p0, caux); \
        else                                                                                     \
            for (; i < end; i += W, a += W * (D * 256)) check_node<D>(a, syn_bit_direct(i), fresh, p0, caux); \
    }
                    BP_DEGREE_SWITCH(
                        deg, BP_CASE, if (BIG) {
                            for (; i < end; i += W) {
                                double *base = reinterpret_cast<double *>(msg_generic + static_cast<size_t>(sb + (i - first) * deg) * 256);
                                check_update_big([&](int k) -> double & { return base[k * 32]; }, deg, syn_bit(i), fresh, p0);
                            }
                        })
#undef BP_CASE
                    if (deg == 0) i += ((end - i + W - 1) / W) * W;   // isolated checks: nothing to send
                }
            } else {
                for (int i = warp; i < p.s; i += W) {
                    const int rp = rowptr_at(i);
                    const int deg = rowptr_at(i + 1) - rp;
                    const bool neg = syn_bit(i);
                    const MH a = ml + rp * 256;
#define BP_CASE(D) check_node<D>(a, neg, fresh, p0, caux)
                    BP_DEGREE_SWITCH(
                        deg, BP_CASE, if (BIG) {
                            double *base = reinterpret_cast<double *>(msg_generic + static_cast<size_t>(rp) * 256);
                            check_update_big([&](int k) -> double & { return base[k * 32]; }, deg, neg, fresh, p0);
                        })
#undef BP_CASE
                }
            }
        }
        __syncthreads();
        // --------------------------------------------------------------- variable pass (:152-178)
        // warp w owns variables w, w+W, ...; decision of its i-th variable = bit i of newbits
        {
            unsigned long long newbits = 0;
            unsigned long long flips = 0;
            uint32_t neww = 0;
            int delta = 0;
            // posterior ratios: every iteration, or (ratio_last_only, the OSD pipeline) only in iteration max_iters --
            // the only ratios belief_propagation_osd.jl:52 ever reads are those of syndromes that did not converge
            const bool wr = p.ratio != nullptr && (!p.ratio_last_only || iter + 1 >= p.max_iters);
            // a lane walks its flipped variables (bits of f, first bit = variable index ibase)
            auto apply_flips = [&](unsigned long long f, int ibase) {
                while (f) {
                    const int j = warp + (ibase + __ffsll(static_cast<long long>(f)) - 1) * W;
                    f &= f - 1;
                    const int e1 = colptr_at(j + 1);
                    for (int e = colptr_at(j); e < e1; ++e) {
                        const uint32_t ent = vflip_at(e);                 // (check/32)*128 + check%32
                        const uint32_t old = atomicXor(resid + (ent >> 7) * 32 + lane, 1u << (ent & 31u));
                        delta += 1 - 2 * static_cast<int>((old >> (ent & 31u)) & 1u);
                    }
                }
            };
            // posterior ratio R of this lane's i-th variable j: optional output, hard decision (:163-168)
            auto record = [&](int j, int i, double R) {
                if (wr) p.ratio[sid * p.n + vorig_at(j)] = R;
                const uint32_t bit = decide(R) ? 1u : 0u;                            // tie -> 1
                if (use_regs) {
                    newbits |= static_cast<unsigned long long>(bit) << i;
                } else {
                    neww |= bit << (i & 31);
                    if ((i & 31) == 31 || j + W >= p.n) {                             // field word complete
                        uint32_t *fw = efield + static_cast<size_t>(i >> 5) * blockDim.x;
                        const uint32_t f = *fw ^ neww;
                        if (f) {
                            *fw = neww;
                            apply_flips(f, i & ~31);
                        }
                        neww = 0;
                    }
                }
            };
            if (staged) {
                if constexpr (!kMsgShared) {
                  if (p.seg.nvseg > 0) {
                    const int nslot = p.pd + 1;
                    const uint32_t ring_c = sbase + p.off_ring + warp * p.ring_warp_bytes + (lane & 15) * 16;
                    const uint32_t ring_l = sbase + p.off_ring + warp * p.ring_warp_bytes + lane * 8;
                    unsigned char *cta_msg = msg_generic - lane * 8 + (lane & 15) * 16;
                    auto off_at = [&](int e) -> uint32_t {
                        if constexpr (kStateShared) return lds_u32(ve_a + 4 * e);
                        else return __ldg(p.g_ve_off + e);
                    };
                    int j = warp, i = 0;
                    for (int g = 0; g < p.seg.nvseg; ++g) {
                        const int deg = p.seg.vdeg[g], first = p.seg.vfirst[g], end = p.seg.vend[g], eb = p.seg.vedge[g];
                        if (deg == 0) {
                            for (; j < end; j += W, ++i)
                                if (active) record(j, i, p0);
                            continue;
                        }
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        int jj = j, ie = eb + (j - first) * D, islot = 0, cslot = 0;                             \
        /* mode 2 reads the slot offsets from global memory (L2 latency): the copy offsets of the next two nodes to be    \
           issued (a0, a1: this half-warp's rows) and the store offsets of the next node to be computed (sn) are kept   \
           in registers, loaded a trip or two before they are needed */                          \
        constexpr bool kVarPipe = BP_M2_VAR_PIPE && !kStateShared;                               \
        constexpr int HD = (D + 1) / 2;                                                          \
        uint32_t a0[HD], a1[HD], sc[D], sn[D];                                                   \
        auto ld_copy = [&](int jn, int en, uint32_t (&o)[HD]) {                                  \
            _Pragma("unroll") for (int k = 0; k < D; k += 2) {                                   \
                o[k / 2] = 0;                                                                    \
                if (jn < end && k + (lane >> 4) < D) o[k / 2] = __ldg(p.g_ve_off + en + k + (lane >> 4)); \
            }                                                                                    \
        };                                                                                       \
        auto ld_store = [&](int jn, int en, uint32_t (&o)[D]) {                                  \
            _Pragma("unroll") for (int k = 0; k < D; ++k) o[k] = jn < end ? __ldg(p.g_ve_off + en + k) : 0u; \
        };                                                                                       \
        if constexpr (kVarPipe) { ld_copy(jj, ie, a0); ld_copy(jj + W, ie + W * D, a1); ld_store(j, ie, sc); } \
        auto issue = [&]() {                                                                     \
            if (jj < end) {                                                                      \
                _Pragma("unroll") for (int k = 0; k < D; k += 2)                                 \
                    if (k + (lane >> 4) < D) {                                                   \
                        uint32_t o_;                                                             \
                        if constexpr (!kVarPipe) o_ = off_at(ie + k + (lane >> 4)); else o_ = a0[k / 2]; \
                        cp_async16(ring_c + islot * p.ring_slot_bytes + (k + (lane >> 4)) * 256, cta_msg + o_); \
                    }                                                                            \
            }                                                                                    \
            cp_async_commit();                                                                   \
            jj += W; ie += W * D;                                                                \
            islot = (islot + 1 == nslot) ? 0 : islot + 1;                                        \
            if constexpr (kVarPipe) {                                                       \
                _Pragma("unroll") for (int k = 0; k < HD; ++k) a0[k] = a1[k];                    \
                ld_copy(jj + W, ie + W * D, a1);                                                 \
            }                                                                                    \
        };                                                                                       \
        TH vea = ve_handle(eb + (j - first) * D);                                                \
        /* Several variables per trip while the ring slot (sized for the widest node, times option ring_mult) holds the   \
           rows of all of them: a fraction of the per-trip bookkeeping (group wait, warp syncs, slot rotation, loop      \
           control) and more bytes in flight per warp */                                         \
        auto multi = [&](auto nv_c) {                                                            \
            constexpr int NV = decltype(nv_c)::value;                                            \
            auto issue_n = [&]() {                                                               \
                if (jj + (NV - 1) * W < end) {                                                   \
                    _Pragma("unroll") for (int k = 0; k < NV * D; k += 2) {                      \
                        const int r = k + (lane >> 4);                                           \
                        if (r < NV * D)                                                          \
                            cp_async16(ring_c + islot * p.ring_slot_bytes + r * 256,             \
                                       cta_msg + off_at(ie + (r / D) * (W * D) + r % D));        \
                    }                                                                            \
                }                                                                                \
                cp_async_commit();                                                               \
                jj += NV * W; ie += NV * W * D;                                                  \
                islot = (islot + 1 == nslot) ? 0 : islot + 1;                                    \
            };                                                                                   \
            const int step_n = W * D * (kStateShared ? 4 : 1);                                   \
            for (int t = 0; t < p.pd; ++t) issue_n();                                            \
            for (; j + (NV - 1) * W < end; j += NV * W, i += NV, vea += NV * step_n) {           \
                issue_n();                                                                       \
                cp_async_wait_pending(p.pd);                                                     \
                __syncwarp();                                                                    \
                if (active) {                                                                    \
                    const uint32_t ra_ = ring_l + cslot * p.ring_slot_bytes;                     \
                    double Rn[NV];                                                               \
                    _Pragma("unroll") for (int v = 0; v < NV; ++v)                               \
                        Rn[v] = var_node_staged<D>(ra_ + v * (D * 256), msg_generic, vea + v * step_n, p0, regular_p0); \
                    _Pragma("unroll") for (int v = 0; v < NV; ++v) record(j + v * W, i + v, Rn[v]); \
                }                                                                                \
                __syncwarp();                                                                    \
                cslot = (cslot + 1 == nslot) ? 0 : cslot + 1;                                    \
            }                                                                                    \
            cp_async_wait_all();                                                                 \
            jj = j; ie = eb + (j - first) * D; islot = 0; cslot = 0;                             \
        };                                                                                       \
        if constexpr (!kVarPipe && D <= 3) {                                                     \
            if (4 * D * 256 <= p.ring_slot_bytes) multi(std::integral_constant<int, 4>());       \
        }                                                                                        \
        if constexpr (!kVarPipe && D <= 6) {                                                     \
            if (2 * D * 256 <= p.ring_slot_bytes) multi(std::integral_constant<int, 2>());       \
        }                                                                                        \
        for (int t = 0; t < p.pd; ++t) issue();                                                  \
        int ec = eb + (j - first) * D;                                                           \
        for (; j < end; j += W, ++i, vea += W * D * (kStateShared ? 4 : 1), ec += W * D) {       \
            issue();                                                                             \
            if constexpr (kVarPipe) ld_store(j + W, ec + W * D, sn);                        \
            cp_async_wait_pending(p.pd);                                                         \
            __syncwarp();                                                                        \
            if (active) {                                                                        \
                if constexpr (!kVarPipe) record(j, i, var_node_staged<D>(ring_l + cslot * p.ring_slot_bytes, msg_generic, vea, p0, regular_p0)); \
                else record(j, i, var_node_staged_off<D>(ring_l + cslot * p.ring_slot_bytes, msg_generic, sc, p0, regular_p0)); \
            }                                                                                    \
            __syncwarp();                                                                        \
            if constexpr (kVarPipe) { _Pragma("unroll") for (int k = 0; k < D; ++k) sc[k] = sn[k]; } \
            cslot = (cslot + 1 == nslot) ? 0 : cslot + 1;                                        \
        }                                                                                        \
        cp_async_wait_all();                                                                     \
    }
                        BP_DEGREE_SWITCH(
                            deg, BP_CASE, if (BIG) {
                                for (; j < end; j += W, ++i) {
                                    if (active) {
                                        const int cp = eb + (j - first) * deg;
                                        record(j, i, var_update_big(
                                            [&](int k) -> double & { return *reinterpret_cast<double *>(msg_generic + off_at(cp + k)); },
                                            deg, p0));
                                    }
                                }
                            })
#undef BP_CASE
                    }
                  } else {
                    const int nslot = p.pd + 1;
                    const uint32_t ring_w = sbase + p.off_ring + warp * p.ring_warp_bytes;
                    unsigned char *cta_msg = msg_generic - lane * 8;
                    auto span = [&](int j, int &cp, int &deg) {
                        if (p.uni_vdeg) { deg = p.uni_vdeg; cp = j * deg; }
                        else { cp = colptr_at(j); deg = colptr_at(j + 1) - cp; }
                    };
                    auto off_at = [&](int e) -> uint32_t {
                        if constexpr (kStateShared) return lds_u32(ve_a + 4 * e);
                        else return __ldg(p.g_ve_off + e);
                    };
                    auto issue = [&](int j, int slot) {
                        if (j < p.n) {
                            int cp, deg;
                            span(j, cp, deg);
                            if (deg <= kMaxRegDegree) {
                                const uint32_t dst = ring_w + slot * p.ring_slot_bytes + (lane & 15) * 16;
                                for (int k = lane >> 4; k < deg; k += 2)
                                    cp_async16(dst + k * 256, cta_msg + off_at(cp + k) + (lane & 15) * 16);
                            }
                        }
                        cp_async_commit();
                    };
                    int islot = 0;
                    for (int t = 0; t < p.pd; ++t) { issue(warp + t * W, islot); islot = (islot + 1 == nslot) ? 0 : islot + 1; }
                    int cslot = 0, i = 0;
                    for (int j = warp; j < p.n; j += W, ++i) {
                        issue(j + p.pd * W, islot);
                        islot = (islot + 1 == nslot) ? 0 : islot + 1;
                        cp_async_wait_pending(p.pd);
                        __syncwarp();
                        if (active) {
                            int cp, deg;
                            span(j, cp, deg);
                            const uint32_t ra = ring_w + cslot * p.ring_slot_bytes + lane * 8;
                            const TH vea = ve_handle(cp);
                            double R = p0;                                            // degree 0: prior only
#define BP_CASE(D) R = var_node_staged<D>(ra, msg_generic, vea, p0, regular_p0)
                            BP_DEGREE_SWITCH(
                                deg, BP_CASE, if (BIG) {
                                    R = var_update_big(
                                        [&](int k) -> double & { return *reinterpret_cast<double *>(msg_generic + off_at(cp + k)); },
                                        deg, p0);
                                })
#undef BP_CASE
                            record(j, i, R);
                        }
                        __syncwarp();
                        cslot = (cslot + 1 == nslot) ? 0 : cslot + 1;
                    }
                    cp_async_wait_all();
                  }
                }
            } else if (active) {
                if (p.uni_vdeg && use_regs) {
                    // two variables per trip for low degrees: their table and message loads are
                    // issued together, halving the exposed shared-memory latency per variable
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        TH vea = ve_handle(warp * D);                                                            \
        int i = 0, j = warp;                                                                     \
        if constexpr (D <= 4) {                                                                  \
            const int step = W * D * (kStateShared ? 4 : 1);                                     \
            for (; j + W < p.n; j += 2 * W, i += 2, vea += 2 * step) {                           \
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
                if (wr) { p.ratio[sid * p.n + j] = Ra; p.ratio[sid * p.n + j + W] = Rb; }   \
                newbits |= static_cast<unsigned long long>((decide(Ra) ? 1u : 0u) | (decide(Rb) ? 2u : 0u)) << i; \
            }                                                                                    \
        }                                                                                        \
        for (; j < p.n; j += W, ++i, vea += W * D * (kStateShared ? 4 : 1)) {                    \
            const double R = var_node<D>(ml, vea, p0, regular_p0);                               \
            if (wr) p.ratio[sid * p.n + j] = R;                                             \
            newbits |= static_cast<unsigned long long>(decide(R) ? 1u : 0u) << i;               \
        }                                                                                        \
    }
                    BP_DEGREE_SWITCH(p.uni_vdeg, BP_CASE, ;)
#undef BP_CASE
                } else if (p.seg.nvseg > 0 && use_regs) {
                    int j = warp, i = 0;
                    for (int g = 0; g < p.seg.nvseg; ++g) {
                        const int deg = p.seg.vdeg[g], first = p.seg.vfirst[g], end = p.seg.vend[g], eb = p.seg.vedge[g];
                        if (deg == 0) {                                               // isolated variables: prior only
                            for (; j < end; j += W, ++i) {
                                if (wr) p.ratio[sid * p.n + vorig_at(j)] = p0;
                                newbits |= static_cast<unsigned long long>(decide(p0) ? 1u : 0u) << i;
                            }
                            continue;
                        }
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        TH vea = ve_handle(eb + (j - first) * D);                                                \
        if constexpr (D <= 4) {                                                                  \
            const int step = W * D * (kStateShared ? 4 : 1);                                     \
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
                if (wr) {                                                                   \
                    p.ratio[sid * p.n + vorig_at(j)] = Ra;                                       \
                    p.ratio[sid * p.n + vorig_at(j + W)] = Rb;                                   \
                }                                                                                \
                newbits |= static_cast<unsigned long long>((decide(Ra) ? 1u : 0u) | (decide(Rb) ? 2u : 0u)) << i; \
            }                                                                                    \
        }                                                                                        \
        for (; j < end; j += W, ++i, vea += W * D * (kStateShared ? 4 : 1)) {                    \
            const double R = var_node<D>(ml, vea, p0, regular_p0);                               \
            if (wr) p.ratio[sid * p.n + vorig_at(j)] = R;                                   \
            newbits |= static_cast<unsigned long long>(decide(R) ? 1u : 0u) << i;               \
        }                                                                                        \
    }
                        BP_DEGREE_SWITCH(
                            deg, BP_CASE, if (BIG) {
                                for (; j < end; j += W, ++i) {
                                    const int cp = eb + (j - first) * deg;
                                    const double R = var_update_big(
                                        [&](int k) -> double & {
                                            uint32_t off;
                                            if constexpr (kStateShared) off = lds_u32(ve_a + 4 * (cp + k));
                                            else off = __ldg(p.g_ve_off + cp + k);
                                            return *reinterpret_cast<double *>(msg_generic + off);
                                        },
                                        deg, p0);
                                    if (wr) p.ratio[sid * p.n + vorig_at(j)] = R;
                                    newbits |= static_cast<unsigned long long>(decide(R) ? 1u : 0u) << i;
                                }
                            })
#undef BP_CASE
                    }
                } else {
                    // general degrees and/or more than 64 variables per warp (decision fields in memory)
                    int i = 0;
                    for (int j = warp; j < p.n; j += W, ++i) {
                        const int cp = colptr_at(j);
                        const int deg = colptr_at(j + 1) - cp;
                        const TH vea = ve_handle(cp);
                        double R = p0;                                                // degree 0: prior only
#define BP_CASE(D) R = var_node<D>(ml, vea, p0, regular_p0)
                        BP_DEGREE_SWITCH(
                            deg, BP_CASE, if (BIG) {
                                R = var_update_big(
                                    [&](int k) -> double & {
                                        uint32_t off;
                                        if constexpr (kStateShared) off = lds_u32(ve_a + 4 * (cp + k));
                                        else off = __ldg(p.g_ve_off + cp + k);
                                        return *reinterpret_cast<double *>(msg_generic + off);
                                    },
                                    deg, p0);
                            })
#undef BP_CASE
                        record(j, i, R);
                    }
                }
            }
            if (active) {
                if (use_regs) {
                    flips = ebits ^ newbits;
                    ebits = newbits;
                }
                // Only variables whose decision flipped touch the residual syndrome s xor H*e
                // (syndrome re-check :180-181, kept incrementally; lanes walk their own flips).
                if (flips) {
                    if (p.uni_vdeg) {
                        uint32_t *rl = resid + lane;
#define BP_CASE(D)                                                                               \
    {                                                                                            \
        delta += flip_walk_uniform<D>(static_cast<uint32_t>(flips), 0, warp, W, vflip_at, rl);   \
        if (flips >> 32) delta += flip_walk_uniform<D>(static_cast<uint32_t>(flips >> 32), 32, warp, W, vflip_at, rl); \
    }
                        BP_DEGREE_SWITCH(p.uni_vdeg, BP_CASE, ;)
#undef BP_CASE
                    } else {
                        apply_flips(flips, 0);
                    }
                }
                if (delta) atomicAdd(nnz + par * 32 + lane, delta);
            }
        }
        if constexpr (kStateShared) {
            if (warp == wP) cp_async_wait_all();              // staged window complete before anyone reads it
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
            auto emit = [&](unsigned long long b, int ibase) {
                while (b) {
                    const int j = vorig_at(warp + (ibase + __ffsll(static_cast<long long>(b)) - 1) * W);
                    b &= b - 1;
                    atomicOr(p.err_words + sid * p.NW + (j >> 5), 1u << (j & 31));
                }
            };
            if (use_regs) emit(ebits, 0);
            else
                for (int k = 0; k < p.nfw; ++k) emit(efield[static_cast<size_t>(k) * blockDim.x], k * 32);
        }
        if (warp == wO && done) {
            p.conv[sid] = conv ? 1 : 0;
            if (p.iters) p.iters[sid] = iter;
            n_done += 1; n_conv += conv ? 1 : 0; n_iters += iter;
        }
        if (warp == wS && !done) nnz[(par ^ 1) * 32 + lane] = cur_nnz;   // carry over to the other buffer
        par ^= 1;
        if (done_mask) refill(done_mask, par);
        __syncthreads();
    }

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
