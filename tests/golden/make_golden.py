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
        if H.shape[1] <= 300:      # dense text twins (small codes only; kept for older readers of this directory)
            np.savetxt(os.path.join(HERE, name + ".H.txt"), H.toarray().astype(int), fmt="%d")
            np.savetxt(os.path.join(HERE, name + ".syndromes.txt"), syn.astype(int), fmt="%d")
            np.savetxt(os.path.join(HERE, name + ".meta.txt"), np.array([per, mi]))
        # text twins for oracle/dump_golden.jl (every case; H as 1-based "row col" pairs): inputs AND the oracle's outputs,
        # so that the Julia script can compare the real package with them and fail loudly
        tdir = os.path.join(HERE, "julia_twins")
        os.makedirs(tdir, exist_ok=True)
        coo = H.tocoo()
        np.savetxt(os.path.join(tdir, name + ".H.coo.txt"), np.stack([coo.row + 1, coo.col + 1], axis=1), fmt="%d")
        np.savetxt(os.path.join(tdir, name + ".meta.txt"), np.array([[H.shape[0], H.shape[1], per, mi]]), fmt="%.17g")
        np.savetxt(os.path.join(tdir, name + ".syndromes.txt"), syn.astype(int), fmt="%d")
        np.savetxt(os.path.join(tdir, name + ".errors.txt"), r["errors"].astype(int), fmt="%d")
        np.savetxt(os.path.join(tdir, name + ".converged.txt"), r["converged"].astype(int), fmt="%d")
        np.savetxt(os.path.join(tdir, name + ".iters.txt"), r["iters"].astype(int), fmt="%d")
        # BP + OSD-0 on the same syndromes with a short iteration budget, so that many of them reach the OSD stage
        mi_osd = 6
        ro = oracle.bposd_decode(H, per, mi_osd, syn)
        np.savetxt(os.path.join(tdir, name + ".osd_meta.txt"), np.array([[mi_osd, int((~ro["converged"]).sum())]]), fmt="%d")
        np.savetxt(os.path.join(tdir, name + ".osd_errors.txt"), ro["errors"].astype(int), fmt="%d")
        np.savetxt(os.path.join(tdir, name + ".osd_converged.txt"), ro["converged"].astype(int), fmt="%d")
        np.savez_compressed(os.path.join(HERE, name + ".osd.npz"), max_iters=mi_osd, errors=ro["errors"], converged=ro["converged"],
                            pivots=ro["pivots"])
        print(name, H.shape, "converged", r["converged"].mean(), "mean iters", r["iters"].mean(), "| OSD case: unconverged",
              int((~ro["converged"]).sum()), "of", B)


if __name__ == "__main__":
    main()
