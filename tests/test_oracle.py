"""CPU tests that pin the oracle (oracle/bp_oracle.c): known answers on trees, an independent
dense transliteration, the reference's semantic edge cases and its statistical thresholds."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import bp_dense

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def random_tree_H(rng, n_checks, max_deg=4):
    """Parity-check matrix whose Tanner graph is a tree (BP is exact there)."""
    rows = []
    n = 1
    frontier = [0]
    for _ in range(n_checks):
        v = frontier[rng.integers(len(frontier))]
        deg = int(rng.integers(2, max_deg + 1))
        new = list(range(n, n + deg - 1))
        n += deg - 1
        rows.append([v] + new)
        frontier += new
    H = np.zeros((n_checks, n), dtype=np.uint8)
    for i, r in enumerate(rows):
        H[i, r] = 1
    return H


def test_tree_kat_survey_example(oracle):
    H = np.array([[1, 1, 0, 0, 0], [0, 1, 1, 1, 0], [0, 0, 0, 1, 1]])
    r = oracle.batch_decode(H, 0.1, 10, np.array([[1], [1], [1]]), want_ratio=True)
    np.testing.assert_allclose(r["ratio"].ravel(), [1, 1, 1 / 9, 1, 1], rtol=1e-12)
    np.testing.assert_allclose(r["ratio"].ravel(), bp_dense.brute_force_ratio(H, 0.1, [1, 1, 1]), rtol=1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_tree_posteriors_equal_enumeration(oracle, seed):
    rng = np.random.default_rng(seed)
    H = random_tree_H(rng, n_checks=int(rng.integers(2, 6)), max_deg=3)
    n = H.shape[1]
    assert n <= 16
    per = float(rng.choice([0.03, 0.1, 0.2]))
    e = (rng.random(n) < 0.3).astype(np.uint8)
    syn = (H @ e) % 2
    # non-early-stopped posterior after >= diameter iterations: ask for ratios at a fixed iteration
    # count by making convergence impossible to stop on (ratios are those of the last iteration run)
    r = oracle.batch_decode(H, per, 40, syn.reshape(-1, 1), want_ratio=True)
    exact = bp_dense.brute_force_ratio(H, per, syn)
    if not r["converged"][0] or r["iters"][0] >= H.shape[0] + 2:
        np.testing.assert_allclose(r["ratio"].ravel(), exact, rtol=1e-9)
    # hard decision of the exact marginals is what a converged tree decode must output
    d, c, ratio, it = bp_dense.decode_dense(H, per, 40, syn)
    assert np.array_equal(d.astype(np.uint8), r["errors"][:, 0])
    assert c == bool(r["converged"][0]) and it == r["iters"][0]


@pytest.mark.parametrize("name,per,B", [("C3", 0.05, 24), ("C2", 0.08, 12), ("C3", 0.15, 12)])
def test_dense_transliteration_equals_c_oracle(oracle, codes, name, per, B):
    H, _, mi = codes.config_matrix(name)
    _, syn = oracle.sample(H, per, 777, 0, B)
    r = oracle.batch_decode(H, per, mi, syn, want_ratio=True)
    Hd = H.toarray()
    for b in range(B):
        e, c, ratio, it = bp_dense.decode_dense(Hd, per, mi, syn[:, b])
        assert np.array_equal(e.astype(np.uint8), r["errors"][:, b])
        assert c == bool(r["converged"][b]) and it == r["iters"][b]
        assert np.array_equal(ratio, r["ratio"][:, b], equal_nan=True)


def test_dense_mode_and_threads_are_bit_identical(oracle, codes):
    H, _, mi = codes.config_matrix("C4")
    _, syn = oracle.sample(H, 0.05, 99, 0, 48)
    a = oracle.batch_decode(H, 0.05, mi, syn, want_ratio=True)
    b = oracle.batch_decode(H, 0.05, mi, syn, want_ratio=True, dense=True)
    c = oracle.batch_decode(H, 0.05, mi, syn, want_ratio=True, nthreads=4)
    for other in (b, c):
        assert np.array_equal(a["errors"], other["errors"])
        assert np.array_equal(a["converged"], other["converged"])
        assert np.array_equal(a["iters"], other["iters"])
        assert np.array_equal(a["ratio"], other["ratio"], equal_nan=True)
    # Inf / NaN are live on this path: the test must actually exercise them
    assert not np.isfinite(a["ratio"]).all() or a["iters"].max() == mi


def test_semantic_edge_cases(oracle, codes):
    H, _, _ = codes.config_matrix("C3")
    s, n = H.shape
    zero = np.zeros((s, 3), dtype=np.uint8)
    # max_iters = 0: loop never runs -> zero error, converged false even for the zero syndrome
    r = oracle.batch_decode(H, 0.01, 0, zero)
    assert not r["errors"].any() and not r["converged"].any() and (r["iters"] == 0).all()
    # zero syndrome converges in iteration 1 with the zero error
    r = oracle.batch_decode(H, 0.01, 5, zero)
    assert not r["errors"].any() and r["converged"].all() and (r["iters"] == 1).all()
    # per >= 0.5: prior ratio >= 1 -> tie/above -> all-ones decision on the first pass
    r = oracle.batch_decode(H, 0.5, 1, zero, want_ratio=True)
    assert r["errors"].all()            # ratio == 1.0 exactly -> err = 1 (tie rule, :164)
    assert (r["ratio"] == 1.0).all()


def test_weight_one_errors(oracle, codes):
    """Single bit-flips: the gross code decodes all of them; on the surface code degenerate
    boundary qubits may stall BP, but whatever is flagged converged must reproduce the syndrome."""
    for name in ("C2", "C3"):
        H, _, mi = codes.config_matrix(name)
        s, n = H.shape
        E = np.eye(n, dtype=np.uint8)
        syn = np.asarray((H @ E) % 2)
        r = oracle.batch_decode(H, 0.01, mi, syn)
        if name == "C3":
            assert r["converged"].all()
            assert np.array_equal(r["errors"], E)
        decoded_syn = np.asarray((H @ r["errors"]) % 2)
        ok = (decoded_syn == syn).all(axis=0)
        assert np.array_equal(ok, r["converged"])


def test_reference_statistical_thresholds(oracle, codes):
    """test/test_bp_decoder.jl:46-51 re-expressed on the restatement: (1000,10,9), p=0.01, 100 iters."""
    H = codes.gallager(1000, 10, 9, seed=7)
    errs, syn = oracle.sample(H, 0.01, 4242, 0, 1100)
    r = oracle.batch_decode(H, 0.01, 100, syn, nthreads=oracle.num_threads())
    exact = (r["errors"] == errs).all(axis=0)
    assert exact[0]                                  # @test test_bp_decoder()
    assert 1 - exact[:100].mean() < 0.005            # batch of 100
    assert 1 - exact[100:1100].mean() < 0.001        # 1000 serial decodes
    assert r["converged"][exact].all()


def test_sampler_is_shard_independent(oracle, codes):
    H, _, _ = codes.config_matrix("C3")
    e_all, s_all = oracle.sample(H, 0.05, 5, 0, 100)
    e_a, s_a = oracle.sample(H, 0.05, 5, 0, 37)
    e_b, s_b = oracle.sample(H, 0.05, 5, 37, 63)
    assert np.array_equal(np.hstack([e_a, e_b]), e_all) and np.array_equal(np.hstack([s_a, s_b]), s_all)
    assert np.array_equal((H @ e_all) % 2, s_all)
    assert abs(e_all.mean() - 0.05) < 0.01


@pytest.mark.parametrize("fixture", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.endswith(".osd.npz")) if os.path.isdir(GOLDEN) else [])
def test_golden_fixtures(oracle, fixture):
    """Regression pins generated by tests/golden/make_golden.py (oracle outputs, NOT reference
    outputs -- Julia cannot run here; see the PARITY UNPINNED note in oracle/bp_oracle.c)."""
    z = np.load(os.path.join(GOLDEN, fixture))
    H = sp.csc_matrix((np.ones(len(z["rowval"]), dtype=np.uint8), z["rowval"], z["colptr"]), shape=tuple(z["shape"]))
    r = oracle.batch_decode(H, float(z["per"]), int(z["max_iters"]), z["syndromes"], want_ratio=True)
    assert np.array_equal(r["errors"], z["errors"])
    assert np.array_equal(r["converged"], z["converged"])
    assert np.array_equal(r["iters"], z["iters"])
    assert np.array_equal(r["ratio"].view(np.uint64), z["ratio_bits"])


def test_minsum_definition_sanity(oracle, codes):
    """The min-sum checker (no reference equivalent): single bit-flips on the gross code decode,
    converged <=> syndrome reproduced, threads == serial."""
    H, _, mi = codes.config_matrix("C3")
    n = H.shape[1]
    E = np.eye(n, dtype=np.uint8)
    syn = np.asarray((H @ E) % 2)
    r = oracle.batch_decode(H, 0.01, mi, syn, variant="minsum")
    assert r["converged"].all() and np.array_equal(r["errors"], E)
    _, syn = oracle.sample(H, 0.06, 3, 0, 400)
    a = oracle.batch_decode(H, 0.06, mi, syn, variant="minsum", want_ratio=True)
    b = oracle.batch_decode(H, 0.06, mi, syn, variant="minsum", want_ratio=True, nthreads=4)
    assert np.array_equal(a["errors"], b["errors"]) and np.array_equal(a["ratio"], b["ratio"], equal_nan=True)
    ok = (np.asarray((H @ a["errors"]) % 2) == syn).all(axis=0)
    assert np.array_equal(ok, a["converged"])


# ---- OSD-0 restatement (belief_propagation_osd.jl:49-125) -------------------------------------
def _gf2_rank(M):
    M = (np.asarray(M) % 2).astype(np.uint8).copy()
    r = 0
    for c in range(M.shape[1]):
        nz = np.nonzero(M[r:, c])[0]
        if nz.size == 0:
            continue
        M[[r, r + nz[0]]] = M[[r + nz[0], r]]
        for i in np.nonzero(M[:, c])[0]:
            if i != r:
                M[i] ^= M[r]
        r += 1
        if r == M.shape[0]:
            break
    return r


@pytest.mark.parametrize("name,per,mi,B", [("C3", 0.08, 8, 80), ("C2", 0.08, 6, 40)])
def test_osd0_c_restatement_equals_dense_transliteration(oracle, codes, name, per, mi, B):
    from oracle import bp_dense
    H, _, _ = codes.config_matrix(name)
    Hd = np.asarray(H.todense()).astype(np.uint8)
    _, syn = oracle.sample(H, per, 31337, 0, B)
    out = oracle.bposd_decode(H, per, mi, syn, nthreads=2)
    bp = oracle.batch_decode(H, per, mi, syn, want_ratio=True)
    assert np.array_equal(out["bp_errors"], bp["errors"]) and np.array_equal(out["converged"], bp["converged"])
    assert (~out["converged"]).sum() > 5
    for b in range(B):
        ref = bp_dense.osd0_dense(Hd, syn[:, b], bp["errors"][:, b], bp["ratio"][:, b])
        assert np.array_equal(ref, out["errors"][:, b]), b
    # converged syndromes come back untouched (:72-74); every output reproduces its syndrome
    cv = out["converged"]
    assert np.array_equal(out["errors"][:, cv], bp["errors"][:, cv])
    assert np.array_equal((Hd.astype(np.int64) @ out["errors"].astype(np.int64)) % 2, syn.astype(np.int64))
    # the pivots never exceed the rank, threads do not change results
    assert out["pivots"].max() <= _gf2_rank(Hd)
    assert np.array_equal(oracle.bposd_decode(H, per, mi, syn, nthreads=1)["errors"], out["errors"])


def test_osd0_reference_testset_properties(oracle, codes):
    """What the reference's own BP+OSD tests pin (test/test_bposd_decoder.jl): the OSD output always
    satisfies the syndrome, also where BP alone fails, and equals BP's output where BP converged."""
    H = codes.gallager(120, 6, 3, seed=11)
    Hd = np.asarray(H.todense()).astype(np.int64)
    _, syn = oracle.sample(H, 0.09, 5, 0, 200)
    out = oracle.bposd_decode(H, 0.09, 10, syn, nthreads=2)
    assert (~out["converged"]).sum() > 10
    assert np.array_equal((Hd @ out["errors"].astype(np.int64)) % 2, syn.astype(np.int64))
    bp_ok = ((Hd @ out["bp_errors"].astype(np.int64)) % 2 == syn).all(axis=0)
    assert np.array_equal(bp_ok, out["converged"])


@pytest.mark.parametrize("name", ["gross_p05", "surface15_p03", "hgp_p05", "gallager1000_p03"])
def test_golden_osd_fixtures_and_julia_twins(oracle, name):
    """BP+OSD-0 fixtures (restated reference, belief_propagation_osd.jl:49-125) and the text twins that
    oracle/dump_golden.jl feeds to the real package: the twins must say exactly what the .npz files say."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    zo = np.load(os.path.join(GOLDEN, name + ".osd.npz"))
    H = sp.csc_matrix((np.ones(len(z["rowval"]), dtype=np.uint8), z["rowval"], z["colptr"]), shape=tuple(z["shape"]))
    r = oracle.bposd_decode(H, float(z["per"]), int(zo["max_iters"]), z["syndromes"])
    assert np.array_equal(r["errors"], zo["errors"]) and np.array_equal(r["converged"], zo["converged"])
    assert np.array_equal(r["pivots"], zo["pivots"])
    t = os.path.join(GOLDEN, "julia_twins", name)
    ij = np.loadtxt(t + ".H.coo.txt", dtype=np.int64).reshape(-1, 2)
    meta = np.loadtxt(t + ".meta.txt")
    Ht = sp.csc_matrix((np.ones(len(ij), dtype=np.uint8), (ij[:, 0] - 1, ij[:, 1] - 1)), shape=(int(meta[0]), int(meta[1])))
    assert (Ht != H).nnz == 0 and meta[2] == float(z["per"]) and int(meta[3]) == int(z["max_iters"])
    ld = lambda suffix: np.loadtxt(t + suffix, dtype=np.int64, ndmin=2)
    assert np.array_equal(ld(".syndromes.txt"), z["syndromes"])
    assert np.array_equal(ld(".errors.txt"), z["errors"]) and np.array_equal(ld(".converged.txt").ravel(), z["converged"].astype(int))
    assert np.array_equal(ld(".iters.txt").ravel(), z["iters"])
    assert np.array_equal(ld(".osd_errors.txt"), zo["errors"]) and np.array_equal(ld(".osd_converged.txt").ravel(), zo["converged"].astype(int))
    assert int(ld(".osd_meta.txt").ravel()[0]) == int(zo["max_iters"])


def test_higher_order_osd_restatement(oracle, codes):
    """osd(..., Val{O}), O > 0 (belief_propagation_osd.jl:127-209): the C restatement equals the independent dense
    transliteration; every output reproduces its syndrome; the weight never exceeds that of the order-0 member of the
    search (x = 0) and does not grow with the order."""
    from oracle import bp_dense
    rng = np.random.default_rng(3)
    cases = [(codes.gross_x(), 0.08, 6), (codes.surface_x(5), 0.08, 4), (codes.gallager(24, 4, 3, seed=2), 0.1, 3)]
    for H, per, mi in cases:
        _, syn = oracle.sample(H, per, 31, 0, 40)
        bp = oracle.batch_decode(H, per, mi, syn, want_ratio=True)
        prev_w = None
        for order in (1, 2, 5):
            r = oracle.bposd_order_decode(H, per, mi, order, syn, nthreads=2)
            assert np.array_equal(r["converged"], bp["converged"])
            for c in range(syn.shape[1]):
                want = bp_dense.osdk_dense(H.toarray(), syn[:, c], bp["errors"][:, c], bp["ratio"][:, c], order)
                assert np.array_equal(r["errors"][:, c], want), (order, c)
            assert ((H @ r["errors"]) % 2 == syn).all()
            w = r["errors"].sum(axis=0)
            if prev_w is not None:
                assert (w <= prev_w).all()
            prev_w = w
    # rank-deficient H with syndromes outside the column space, order larger than the information set
    H = sp.csc_matrix(np.array([[1, 1, 0, 0], [0, 1, 1, 0], [1, 0, 1, 0]], dtype=np.uint8))
    syn = np.array([[1, 0, 1], [1, 1, 1], [0, 0, 1]], dtype=np.uint8).T
    bp = oracle.batch_decode(H, 0.1, 3, syn, want_ratio=True)
    r = oracle.bposd_order_decode(H, 0.1, 3, 7, syn)
    for c in range(syn.shape[1]):
        assert np.array_equal(r["errors"][:, c], bp_dense.osdk_dense(H.toarray(), syn[:, c], bp["errors"][:, c], bp["ratio"][:, c], 7))
