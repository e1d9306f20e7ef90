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
