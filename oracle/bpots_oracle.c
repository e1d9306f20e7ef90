/*
 * bpots_oracle.c -- CPU restatement of LDPCDecoders.jl's BP-OTS decoder.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/bp_oracle.c): nothing in ldpcdecoders.jl_b200/ may include, link or call it.
 * PARITY UNPINNED, and here for one reason more: the reference evaluates tanh, atanh and log with Julia's own
 * implementations, this file with the C library's, the CUDA kernel with CUDA's -- all within an ulp or two of one
 * another but not bit-identical, and BP amplifies that on hard syndromes.  The GPU-vs-this comparison in the tests is
 * therefore statistical for the arithmetic and exact for everything discrete that follows from equal messages.
 *
 * Restated, statement by statement (citations relative to /root/reference/src/decoders/bpots_decoder.jl):
 *   :144-156  reset!                       (messages, oscillation counters, prior decisions cleared)
 *   :161-174  update_variable_to_check!    nu_{j->i} = Omega_j + sum_{i' != i} mu_{i'->j}   (sum from 0.0 in neighbour order)
 *   :180-211  update_check_to_variable!    clamped tanh product over j' != j, sign by the syndrome bit, 2 atanh, clamp +-100
 *   :113-129  compute_beliefs!             llr_j = Omega_j + sum_i mu_{i->j}; decision = llr < 0
 *   :226-340  decode!                      priors log((1 - 2p/3)/(2p/3)), oscillation counting, best-so-far by
 *                                          (mismatch, weight), early return on mismatch 0, biasing every T iterations
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t s, n, E;
    const int64_t *colptr, *rowval;   /* CSC of H, 0-based; var_neighbors[j] = rowval[colptr[j] .. colptr[j+1]) ascending */
    int64_t *rowptr, *rowvar, *rowedge;   /* check_neighbors[i] ascending, and the CSC position of each (i, j) */
} ots_graph_t;

static int ots_build(ots_graph_t *g, int64_t s, int64_t n, const int64_t *colptr, const int64_t *rowval)
{
    g->s = s; g->n = n; g->E = colptr[n]; g->colptr = colptr; g->rowval = rowval;
    g->rowptr = (int64_t *)calloc((size_t)s + 1, sizeof(int64_t));
    g->rowvar = (int64_t *)malloc(sizeof(int64_t) * (size_t)(g->E ? g->E : 1));
    g->rowedge = (int64_t *)malloc(sizeof(int64_t) * (size_t)(g->E ? g->E : 1));
    if (!g->rowptr || !g->rowvar || !g->rowedge) return -1;
    for (int64_t e = 0; e < g->E; ++e) g->rowptr[rowval[e] + 1]++;
    for (int64_t i = 0; i < s; ++i) g->rowptr[i + 1] += g->rowptr[i];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(s ? s : 1));
    memcpy(fill, g->rowptr, sizeof(int64_t) * (size_t)s);
    for (int64_t j = 0; j < n; ++j)              /* :97-108: check_neighbors filled in ascending variable order */
        for (int64_t e = colptr[j]; e < colptr[j + 1]; ++e) {
            const int64_t i = rowval[e];
            g->rowvar[fill[i]] = j; g->rowedge[fill[i]] = e; fill[i]++;
        }
    free(fill);
    return 0;
}

/* One decode!.  vc / cv: E doubles each, indexed by CSC edge.  Returns converged; out = best_decisions (n bytes). */
static int ots_decode(const ots_graph_t *g, double per, int max_iters, int T, double C, const uint8_t *syn, double *vc, double *cv,
                      double *Omega, double *llr, int *osc, uint8_t *dec, uint8_t *prior, uint8_t *best, int32_t *iters_out)
{
    const int64_t s = g->s, n = g->n;
    const double MAX_TANH = 0.99999, MAX_MSG = 100.0;
    for (int64_t e = 0; e < g->E; ++e) { vc[e] = 0.0; cv[e] = 0.0; }        /* reset! */
    for (int64_t j = 0; j < n; ++j) { osc[j] = 0; prior[j] = 0; best[j] = 0; }
    const double Pi = log((1 - (2 * per / 3)) / (2 * per / 3));             /* :231 */
    for (int64_t j = 0; j < n; ++j) Omega[j] = Pi;
    int64_t best_mismatch = s, best_weight = n;                              /* :236-237 (length(syndrome), n) */
    int32_t it = 0;
    for (int iter = 1; iter <= max_iters; ++iter) {
        it = iter;
        for (int64_t j = 0; j < n; ++j)                                      /* :241-245 */
            for (int64_t e = g->colptr[j]; e < g->colptr[j + 1]; ++e) {
                double msg_sum = 0.0;
                for (int64_t e2 = g->colptr[j]; e2 < g->colptr[j + 1]; ++e2)
                    if (e2 != e) msg_sum += cv[e2];
                vc[e] = Omega[j] + msg_sum;
            }
        for (int64_t i = 0; i < s; ++i)                                      /* :247-251 */
            for (int64_t k = g->rowptr[i]; k < g->rowptr[i + 1]; ++k) {
                double prod = 1.0;
                for (int64_t k2 = g->rowptr[i]; k2 < g->rowptr[i + 1]; ++k2)
                    if (k2 != k) {
                        double t = tanh(0.5 * vc[g->rowedge[k2]]);
                        t = fmin(MAX_TANH, fmax(-MAX_TANH, t));
                        prod *= t;
                    }
                if (syn[i]) prod = -prod;
                if (fabs(prod) >= MAX_TANH) prod = prod > 0 ? MAX_TANH : -MAX_TANH;
                double msg = 2.0 * atanh(prod);
                msg = fmin(MAX_MSG, fmax(-MAX_MSG, msg));
                cv[g->rowedge[k]] = msg;
            }
        for (int64_t j = 0; j < n; ++j) {                                    /* compute_beliefs! */
            double l = Omega[j];
            for (int64_t e = g->colptr[j]; e < g->colptr[j + 1]; ++e) l += cv[e];
            llr[j] = l;
            dec[j] = l < 0.0 ? 1 : 0;
        }
        if (iter > 1) for (int64_t j = 0; j < n; ++j) osc[j] += dec[j] ^ prior[j];   /* :257-261 */
        memcpy(prior, dec, (size_t)n);
        int64_t mismatch = 0, weight = 0;                                    /* :265-281 */
        for (int64_t i = 0; i < s; ++i) {
            unsigned par = 0;
            for (int64_t k = g->rowptr[i]; k < g->rowptr[i + 1]; ++k) par ^= dec[g->rowvar[k]];
            mismatch += par != (unsigned)(syn[i] ? 1 : 0);
        }
        for (int64_t j = 0; j < n; ++j) weight += dec[j];
        if (mismatch < best_mismatch || (mismatch == best_mismatch && weight < best_weight)) {   /* :283-292 */
            best_mismatch = mismatch; best_weight = weight;
            memcpy(best, dec, (size_t)n);
            if (mismatch == 0) { if (iters_out) *iters_out = it; return 1; }
        }
        if (mismatch > 0 && iter % T == 0) {                                 /* :294-336 */
            for (int64_t j = 0; j < n; ++j) Omega[j] = Pi;
            int mx = 0;
            for (int64_t j = 0; j < n; ++j) mx = osc[j] > mx ? osc[j] : mx;
            if (mx > 0) {
                int max_osc = 0; int64_t j1 = -1; double min_llr = INFINITY;
                for (int64_t j = 0; j < n; ++j) {
                    if (osc[j] > max_osc) { max_osc = osc[j]; j1 = j; min_llr = fabs(llr[j]); }
                    else if (osc[j] == max_osc && fabs(llr[j]) < min_llr) { j1 = j; min_llr = fabs(llr[j]); }
                }
                if (j1 >= 0) { osc[j1] = 0; Omega[j1] = -C; }
                int64_t j2 = 0; min_llr = fabs(llr[0]);
                for (int64_t j = 1; j < n; ++j) if (fabs(llr[j]) < min_llr) { j2 = j; min_llr = fabs(llr[j]); }
                Omega[j2] = -C;
            }
        }
    }
    if (iters_out) *iters_out = it;
    return 0;
}

/* batchdecode! through the generic column loop (abstract_decoder.jl:31-42).  syn: s x B bytes column-major; err: n x B bytes out. */
int bpots_oracle_batch(int64_t s, int64_t n, const int64_t *colptr, const int64_t *rowval, double per, int32_t max_iters, int32_t T,
                       double C, int64_t B, const uint8_t *syn, uint8_t *err, uint8_t *conv, int32_t *iters, int32_t nthreads)
{
    ots_graph_t g;
    if (ots_build(&g, s, n, colptr, rowval)) return -1;
    if (nthreads < 1) nthreads = 1;
    int fail = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        const size_t E = (size_t)(g.E ? g.E : 1), nn = (size_t)(n ? n : 1);
        double *vc = (double *)malloc(8 * E), *cv = (double *)malloc(8 * E), *Om = (double *)malloc(8 * nn), *llr = (double *)malloc(8 * nn);
        int *osc = (int *)malloc(sizeof(int) * nn);
        uint8_t *dec = (uint8_t *)malloc(nn), *prior = (uint8_t *)malloc(nn);
        if (!vc || !cv || !Om || !llr || !osc || !dec || !prior) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            fail = 1;
        } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
            for (int64_t c = 0; c < B; ++c) {
                int32_t it = 0;
                conv[c] = (uint8_t)ots_decode(&g, per, max_iters, T, C, syn + (size_t)c * (size_t)s, vc, cv, Om, llr, osc, dec, prior,
                                              err + (size_t)c * (size_t)n, &it);
                if (iters) iters[c] = it;
            }
        }
        free(vc); free(cv); free(Om); free(llr); free(osc); free(dec); free(prior);
    }
    free(g.rowptr); free(g.rowvar); free(g.rowedge);
    return fail ? -1 : 0;
}
