"""Small decodes through every kernel mode, meant to run under compute-sanitizer
(memcheck / racecheck) on the GPU box:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
oracle = entry.load_oracle()
rng = np.random.default_rng(3)


def check(H, per, mi, B, **opts):
    _, syn = oracle.sample(H, per, 5, 0, B)
    ref = oracle.batch_decode(H, per, mi, syn)
    dec = pkg.BeliefPropagationDecoder(H, per, mi, **opts)
    err = np.zeros((H.shape[1], B), dtype=np.uint8, order="F")
    it = np.zeros(B, dtype=np.int32)
    _, ok = pkg.batchdecode_b(dec, syn, err, iters=it)
    info = dec.info()
    dec.close()
    bad = int((err != ref["errors"]).any(axis=0).sum() + (ok != ref["converged"]).sum() + (it != ref["iters"]).sum())
    print("mode", info["kernel_mode"], "threads", info["threads_per_cta"], "pd", info["prefetch_distance"], "B", B, "bad", bad, flush=True)
    assert bad == 0


codes = pkg.codes
check(codes.gross_x(), 0.05, 8, 200)                                   # mode 0, uniform degrees
check(codes.surface_x(7), 0.05, 8, 150)                                # mode 0, degree segments
Hc = codes.hgp_x(codes.gallager(16, 4, 3, seed=1))
check(Hc, 0.03, 6, 100)                                                # mode 1, staged
check(Hc, 0.03, 6, 100, prefetch=0)                                    # mode 1, direct
check(codes.gallager(60000, 6, 3, seed=2), 0.02, 4, 40)                # mode 2, staged
H = (rng.random((30, 70)) < 0.08).astype(np.uint8); H[4, 10:40] = 1
check(sp.csc_matrix(H), 0.05, 6, 70)                                   # local-memory degree path
print("sanitize smoke ok")
