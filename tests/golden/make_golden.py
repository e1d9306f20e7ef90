"""Generate the committed regression fixtures (tests/golden/*.npz and the .txt twins that
oracle/dump_golden.jl feeds to the real Julia package).

The expected outputs come from the CPU ORACLE, not from the reference: Julia cannot run in this
image (see the PARITY UNPINNED note in oracle/bp_oracle.c).  They pin the oracle against silent
regressions and give anyone with Julia a ready-made input set to cross-check it.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    pkg = entry.load_package()
    oracle = entry.load_oracle()
    cases = [("gross_p05", "C3", 0.05, 32, 64), ("surface15_p03", "C2", 0.03, 32, 64),
             ("hgp_p05", "C4", 0.05, 32, 16), ("gallager1000_p03", "C1", 0.03, 25, 8)]
    for name, cfg, per, mi, B in cases:
        H, _, _ = pkg.codes.config_matrix(cfg)
        _, syn = oracle.sample(H, per, 20240, 0, B)
        r = oracle.batch_decode(H, per, mi, syn, want_ratio=True)
        Hc = H.tocsc()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array(H.shape), colptr=Hc.indptr.astype(np.int64),
                            rowval=Hc.indices.astype(np.int64), per=per, max_iters=mi, syndromes=syn,
                            errors=r["errors"], converged=r["converged"], iters=r["iters"],
                            ratio_bits=r["ratio"].view(np.uint64))
        if H.shape[1] <= 300:      # text twins for the Julia cross-check (small codes only)
            np.savetxt(os.path.join(HERE, name + ".H.txt"), H.toarray().astype(int), fmt="%d")
            np.savetxt(os.path.join(HERE, name + ".syndromes.txt"), syn.astype(int), fmt="%d")
            np.savetxt(os.path.join(HERE, name + ".meta.txt"), np.array([per, mi]))
        print(name, H.shape, "converged", r["converged"].mean(), "mean iters", r["iters"].mean())


if __name__ == "__main__":
    main()
