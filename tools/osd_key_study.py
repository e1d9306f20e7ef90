#!/usr/bin/env python
"""How often does the OSD-0 sort key 1/R (kernel and oracle) order columns differently from exp(log(1/R)) (the reference's
expression, belief_propagation_osd.jl:53 with belief_propagation.jl:163)?  CPU only.  Two measurements:
  (a) end to end: BP+OSD-0 outputs of the restated reference under both keys on the unconverged syndromes of C3 / C4 / C2;
  (b) key level: posterior ratios of those syndromes -- pairs of columns whose keys are equal under one form and different
      under the other (the only way the stable sort can change), plus a synthetic near-tie class (R and R*(1 +- k ulp)).
Writes a text report (profiles/r2_osd_key_study.txt)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def keys(R, mode):
    with np.errstate(all="ignore"):
        r = 1.0 / R
        if mode == 1:
            r = np.exp(np.log(r))
        return np.maximum(r, 1.0 - r)


def main():
    pkg = entry.load_package()
    o = entry.load_oracle()
    o.build()
    out = []
    for name, per, mi, B in (("C3", 0.08, 32, 20000), ("C3", 0.12, 32, 10000), ("C2", 0.05, 32, 10000), ("C4", 0.04, 32, 1500), ("C1", 0.04, 25, 300)):
        H, _, _ = pkg.codes.config_matrix(name)
        _, syn = o.sample(H, per, 2024, 0, B)
        nt = o.num_threads()
        a = o.bposd_decode(H, per, mi, syn, nthreads=nt, key_mode=0)
        b = o.bposd_decode(H, per, mi, syn, nthreads=nt, key_mode=1)
        unconv = ~a["converged"]
        differ = (a["errors"] != b["errors"]).any(axis=0)
        r = o.batch_decode(H, per, mi, syn[:, unconv], nthreads=nt, want_ratio=True)["ratio"]
        k0, k1 = keys(r, 0), keys(r, 1)
        # columns (per syndrome) involved in an order change: sort both ways (stable, descending) and compare permutations
        changed = 0
        tie_pairs_created = tie_pairs_broken = 0
        for c in range(r.shape[1]):
            p0 = np.argsort(-k0[:, c], kind="stable")
            p1 = np.argsort(-k1[:, c], kind="stable")
            changed += int((p0 != p1).any())
            s0 = np.sort(k0[:, c]); s1 = np.sort(k1[:, c])
            tie_pairs_created += int(((np.diff(s1) == 0) & ~(np.diff(np.sort(k0[:, c])) == 0)).sum())
            tie_pairs_broken += int(((np.diff(s0) == 0) & ~(np.diff(np.sort(k1[:, c])) == 0)).sum())
        out.append("%s per=%.2f max_iters=%d: %d syndromes, %d unconverged; sort permutation differs on %d of them (%.2f %%); "
                   "OSD-0 OUTPUT differs on %d (%.3f %% of the unconverged, %.4f %% of all)" % (
                       name, per, mi, B, int(unconv.sum()), changed, 100.0 * changed / max(int(unconv.sum()), 1),
                       int(differ.sum()), 100.0 * differ.sum() / max(int(unconv.sum()), 1), 100.0 * differ.sum() / B))
    # synthetic near ties: R and its neighbours k ulps away
    rng = np.random.default_rng(7)
    R = np.exp(rng.uniform(-30, 30, 2_000_000))
    for ulps in (1, 2, 4, 16, 256, 65536):
        R2 = (R.view(np.int64) + ulps).view(np.float64)
        a0, b0 = keys(R, 0), keys(R2, 0)
        a1, b1 = keys(R, 1), keys(R2, 1)
        flips = ((a0 < b0) != (a1 < b1)) | ((a0 == b0) != (a1 == b1))
        out.append("near ties: R vs R + %d ulp, 2e6 draws log-uniform in [e^-30, e^30]: relation of the two keys differs between the forms for %.3f %%" % (
            ulps, 100.0 * flips.mean()))
    text = "\n".join(out)
    print(text)
    with open(os.path.join(ROOT, "profiles", "r2_osd_key_study.txt"), "w") as f:
        f.write("# tools/osd_key_study.py -- sort key RN(1/R) (kernel, oracle) against exp(log(1/R)) with glibc (the reference's expression; Julia's own\n"
                "# exp/log differ from glibc's in the last bit, so this measures the KIND of deviation, not Julia's exact behaviour)\n" + text + "\n")


if __name__ == "__main__":
    main()
