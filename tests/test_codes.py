import numpy as np


def test_config_shapes(codes):
    expect = {"C1": (900, 1000, 9000), "C2": (112, 225, 420), "C3": (72, 144, 432),
              "C4": (768, 1600, 5376), "C5": (50001, 100002, 300006)}
    for name, (s, n, E) in expect.items():
        H, per, mi = codes.config_matrix(name)
        assert H.shape == (s, n) and H.nnz == E
        assert H.has_sorted_indices


def test_gallager_is_regular(codes):
    """test/test_oldtests.jl:1-17: every row sums to wr, every column to wc."""
    H = codes.gallager(1000, 10, 9, seed=3)
    assert (np.asarray(H.sum(1)).ravel() == 10).all()
    assert (np.asarray(H.sum(0)).ravel() == 9).all()
    try:
        codes.gallager(100000, 6, 3)
        assert False
    except AssertionError:
        pass


def test_css_commutation(codes):
    hx, hz = codes.gross_x().astype(int), codes.gross_z().astype(int)
    assert ((hx @ hz.T).toarray() % 2 == 0).all()


def test_surface_weights(codes):
    H = codes.surface_x(15)
    rs = np.asarray(H.sum(1)).ravel()
    assert sorted(np.unique(rs)) == [2, 4] and (rs == 2).sum() == 14 and (rs == 4).sum() == 98


def test_css_logicals_of_surface_and_gross_codes(codes):
    """Logical operators used for failure counting: commute with the other check type, independent of the stabilizers,
    and as many as the code has logical qubits (surface: 1, gross: 12)."""
    import numpy as np
    for Hx, Hz, k in ((codes.surface_x(5), codes.surface_z(5), 1), (codes.gross_x(), codes.gross_z(), 12)):
        assert ((Hx @ Hz.T).toarray() % 2 == 0).all()
        L = codes.css_logicals(Hx, Hz)
        assert L.shape == (k, Hx.shape[1])
        assert ((Hz @ L.T).toarray() % 2 == 0).all()
        # a stabilizer (row of Hz) is harmless, a logical operator of the other type is not
        assert ((L @ Hz.T).toarray() % 2 == 0).all()
        Lz = codes.css_logicals(Hz, Hx)
        assert np.linalg.matrix_rank((L @ Lz.T).toarray() % 2) >= 1


def test_pcm_text_io_round_trip(codes, tmp_path):
    """save_pcm / load_pcm (parity_generator.jl:47-54): tab-separated 0/1 rows as Julia's writedlm writes them."""
    import numpy as np
    H = codes.parity_check_matrix(60, 6, 3, seed=5)
    assert (np.asarray(H.sum(axis=1)).ravel() == 6).all() and (np.asarray(H.sum(axis=0)).ravel() == 3).all()
    path = tmp_path / "H.txt"
    codes.save_pcm(H, path)
    first = open(path).readline().rstrip("\n").split("\t")
    assert len(first) == 60 and set(first) <= {"0", "1"}
    assert (codes.load_pcm(path) != H).nnz == 0
    (tmp_path / "J.txt").write_text("1.0 0.0 1.0\n0 1 1\n")           # readdlm-style floats
    assert codes.load_pcm(tmp_path / "J.txt").toarray().tolist() == [[1, 0, 1], [0, 1, 1]]
