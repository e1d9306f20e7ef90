#!/usr/bin/env python
"""Randomised parity sweep (not part of the test suite: a few minutes of GPU time): random sparse parity-check matrices of
varied size / degree profile, several error rates and iteration caps, every decode path -- persistent kernel (both families),
node-parallel small-batch kernel, min-sum, BP -> OSD-0 -- bit-compared with the oracle.  Prints one line per case and a
summary; exits non-zero on the first mismatch."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); oracle = entry.load_oracle()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
rng = np.random.default_rng(seed)
nth = oracle.num_threads()


def random_code():
    kind = rng.integers(0, 4)
    if kind == 0:      # column-regular
        n = int(rng.integers(8, 400)); s = int(rng.integers(4, max(5, n)))
        wc = int(rng.integers(1, min(6, s) + 1))
        H = np.zeros((s, n), np.uint8)
        for j in range(n):
            H[rng.choice(s, wc, replace=False), j] = 1
    elif kind == 1:    # Bernoulli entries, possibly empty rows/columns and heavy nodes
        n = int(rng.integers(5, 300)); s = int(rng.integers(3, 200))
        H = (rng.random((s, n)) < rng.uniform(0.01, 0.12)).astype(np.uint8)
    elif kind == 2:    # Gallager ensemble
        wr = int(rng.integers(3, 9)); wc = int(rng.integers(2, 5)); n = wr * int(rng.integers(4, 60))
        H = np.asarray(sp.csc_matrix(pkg.codes.gallager(n, wr, wc, seed=int(rng.integers(1 << 30)))).todense()).astype(np.uint8)
    else:              # hypergraph product of a small random code
        wr, wc = 4, 3
        Hc = pkg.codes.gallager(wr * int(rng.integers(2, 6)), wr, wc, seed=int(rng.integers(1 << 30)))
        H = np.asarray(sp.csc_matrix(pkg.codes.hgp_x(Hc)).todense()).astype(np.uint8) if hasattr(pkg.codes, "hgp_x") else None
    return H


def gpu(H, per, mi, syn, cls=None, **opts):
    dec = (cls or pkg.BeliefPropagationDecoder)(H, per, mi, **opts)
    n, B = H.shape[1], syn.shape[1]
    errors = np.zeros((n, B), dtype=np.uint8, order="F")
    if cls is None:
        iters = np.zeros(B, dtype=np.int32)
        ratio = np.zeros((n, B), dtype=np.float64, order="F") if mi > 0 else None
        _, ok = pkg.batchdecode_b(dec, np.asfortranarray(syn), errors, iters=iters, posterior_ratio=ratio)
        out = (errors, ok.copy(), iters, ratio)
    else:
        _, ok = pkg.batchdecode_b(dec, np.asfortranarray(syn), errors)
        out = (errors, ok.copy(), None, None)
    dec.close()
    return out


t_end = time.time() + budget
cases = 0
while time.time() < t_end:
    H = random_code()
    if H is None or H.sum() == 0 or max(int(H.sum(0).max()), int(H.sum(1).max())) > 128:     # LDPCB200_MAX_DEGREE
        continue
    s, n = H.shape
    Hs = sp.csc_matrix(H)
    per = float(rng.choice([0.002, 0.02, 0.06, 0.15, 0.5]))
    mi = int(rng.choice([0, 1, 2, 7, 20]))
    B = int(rng.choice([1, 3, 40, 149, 700]))
    e = (rng.random((n, B)) < min(per, 0.3)).astype(np.uint8)
    syn = np.asarray((Hs @ e) % 2).astype(np.uint8)
    if rng.random() < 0.3:
        syn ^= (rng.random(syn.shape) < 0.05).astype(np.uint8)        # also syndromes outside the column space
    ref = oracle.batch_decode(Hs, per, mi, syn, nthreads=nth, want_ratio=True)
    maxdeg = max(int(H.sum(0).max()), int(H.sum(1).max()))
    paths = [("auto", {}), ("persistent", dict(small_batch=0)), ("global", dict(family=2, small_batch=0))]
    for name, o in paths:
        g = gpu(Hs, per, mi, syn, **o)
        bad = (g[0] != ref["errors"]).any() or (g[1] != ref["converged"]).any() or (g[2] != ref["iters"]).any() or \
            (mi > 0 and not np.array_equal(g[3].view(np.uint64), ref["ratio"].view(np.uint64)))
        if bad:
            print("MISMATCH", name, "seed", seed, "case", cases, "shape", H.shape, "per", per, "mi", mi, "B", B, "maxdeg", maxdeg)
            np.savez("gpurun_out/fuzz_fail_%d_%d.npz" % (seed, cases), H=H, syn=syn, per=per, mi=mi)
            sys.exit(1)
    extra = ""
    if maxdeg <= 12 and 0 < per < 1:
        refm = oracle.batch_decode(Hs, per, mi, syn, nthreads=nth, variant="minsum")
        for o in ({}, dict(small_batch=0)):
            d = pkg.BeliefPropagationDecoder(Hs, per, mi, variant="minsum", **o)
            em = np.zeros((n, B), dtype=np.uint8, order="F")
            _, ok = pkg.batchdecode_b(d, np.asfortranarray(syn), em)
            d.close()
            if (em != refm["errors"]).any() or (ok != refm["converged"]).any():
                print("MISMATCH minsum", o, "seed", seed, "case", cases, H.shape, per, mi, B)
                np.savez("gpurun_out/fuzz_fail_ms_%d_%d.npz" % (seed, cases), H=H, syn=syn, per=per, mi=mi)
                sys.exit(1)
        extra += " minsum"
    if s <= 2048 and s * ((n + 128) // 128 * 16 + 16) + 16 * n < 200_000:
        refo = oracle.bposd_decode(Hs, per, mi, syn, nthreads=nth)
        go = gpu(Hs, per, mi, syn, cls=pkg.BeliefPropagationOSDDecoder)
        if (go[0] != refo["errors"]).any() or (go[1] != refo["converged"]).any():
            print("MISMATCH osd seed", seed, "case", cases, H.shape, per, mi, B)
            np.savez("gpurun_out/fuzz_fail_osd_%d_%d.npz" % (seed, cases), H=H, syn=syn, per=per, mi=mi)
            sys.exit(1)
        extra += " osd"
    cases += 1
    print("ok case %d: %dx%d maxdeg %d per %.3f max_iters %d B %d conv %.2f%s" % (cases, s, n, maxdeg, per, mi, B, ref["converged"].mean(), extra), flush=True)
print("fuzz: %d random cases, all paths bit-identical to the oracle (seed %d)" % (cases, seed))
