"""Parity-check matrices of the BASELINE.json configurations (harness inputs, seeded).

All functions return ``scipy.sparse.csc_matrix`` with uint8 data and sorted indices, which is
the shape ``SparseMatrixCSC`` has on the reference side (rows ascending inside a column).

* ``gallager``  -- same ensemble as /root/reference/src/parity_generator.jl:21-45 (block of
  ``wr`` consecutive ones per row, ``wc-1`` further column-shuffled copies), built sparsely
  and from a seeded generator (the reference uses Julia's unseeded global RNG, so ``H`` is an
  input to both sides, never something to reproduce bit-for-bit).
* ``surface_x`` -- X-type checks of the distance-d rotated surface code (config C2).
* ``gross_x``   -- H_X = [A|B] of the [[144,12,12]] bivariate-bicycle code (config C3).
* ``hgp_x``     -- H_X of the hypergraph product of a classical code with itself (config C4).
"""
import numpy as np
import scipy.sparse as sp


def _csc(rows, cols, shape):
    data = np.ones(len(rows), dtype=np.uint8)
    H = sp.coo_matrix((data, (np.asarray(rows), np.asarray(cols))), shape=shape).tocsc()
    H.data[:] = H.data % 2
    H.eliminate_zeros()
    H.sort_indices()
    return H


def gallager(n, wr, wc, seed=0):
    """Regular (wc, wr) Gallager matrix, (n*wc/wr) x n."""
    if n % wr != 0:
        raise AssertionError("n % wr == 0 required (parity_generator.jl:25)")
    block = n // wr
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = [np.repeat(np.arange(block), wr)]
    cols = [np.arange(n)]
    for b in range(1, wc):
        perm = rng.permutation(n)          # new column c shows old column perm[c]
        rows.append(b * block + perm // wr)
        cols.append(np.arange(n))
    return _csc(np.concatenate(rows), np.concatenate(cols), (block * wc, n))


def parity_check_matrix(n, wr, wc, seed=0):
    """parity_check_matrix(n, wr, wc) of /root/reference/src/parity_generator.jl:21-45 (same ensemble, seeded)."""
    return gallager(n, wr, wc, seed=seed)


def save_pcm(H, file_path):
    """save_pcm(H, file_path) (parity_generator.jl:47-49): the matrix as 0/1 integers, one row per line, tab separated --
    the text Julia's writedlm(file_path, Int.(H)) produces, so files travel between the two packages."""
    A = np.asarray(sp.csc_matrix(H).todense()).astype(np.int64) & 1
    with open(file_path, "w") as f:
        for row in A:
            f.write("\t".join("1" if x else "0" for x in row) + "\n")


def load_pcm(file_path):
    """load_pcm(file_path) (parity_generator.jl:51-54): reads what save_pcm / Julia's writedlm wrote (any whitespace or comma
    delimiter, integer or float spelling of 0/1) and returns it as a sparse 0/1 matrix (the reference returns Int.(H))."""
    rows = []
    with open(file_path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            rows.append([int(float(tok)) for tok in line.replace(",", " ").split()])
    A = np.array(rows, dtype=np.int64)
    if A.ndim != 2:
        raise ValueError("ragged parity-check file")
    if ((A != 0) & (A != 1)).any():
        raise ValueError("parity-check entries must be 0 or 1")
    return sp.csc_matrix(A.astype(np.uint8))


def surface_x(d):
    """X checks of the rotated surface code: ((d*d-1)/2) x (d*d); bulk weight 4, boundary 2."""
    rows, cols = [], []
    r_idx = 0
    for r in range(d + 1):
        for c in range(d + 1):
            if (r + c) % 2 != 0:
                continue                   # Z-type plaquette
            qs = [(rr, cc) for rr in (r - 1, r) for cc in (c - 1, c) if 0 <= rr < d and 0 <= cc < d]
            bulk = 1 <= r <= d - 1 and 1 <= c <= d - 1
            top_bottom = (r == 0 or r == d) and 1 <= c <= d - 1
            if not (bulk or top_bottom):
                continue
            for (rr, cc) in qs:
                rows.append(r_idx)
                cols.append(rr * d + cc)
            r_idx += 1
    return _csc(rows, cols, (r_idx, d * d))


def surface_z(d):
    """Z checks of the same rotated surface code (the other colour of plaquettes; left/right boundary weight 2)."""
    rows, cols = [], []
    r_idx = 0
    for r in range(d + 1):
        for c in range(d + 1):
            if (r + c) % 2 != 1:
                continue                   # X-type plaquette
            qs = [(rr, cc) for rr in (r - 1, r) for cc in (c - 1, c) if 0 <= rr < d and 0 <= cc < d]
            bulk = 1 <= r <= d - 1 and 1 <= c <= d - 1
            left_right = (c == 0 or c == d) and 1 <= r <= d - 1
            if not (bulk or left_right):
                continue
            for (rr, cc) in qs:
                rows.append(r_idx)
                cols.append(rr * d + cc)
            r_idx += 1
    return _csc(rows, cols, (r_idx, d * d))


def _gf2_rref(M):
    """Reduced row echelon form over GF(2) of a dense 0/1 array; returns (R, pivot columns)."""
    R = (np.array(M, dtype=np.uint8) & 1).copy()
    piv = []
    r = 0
    for c in range(R.shape[1]):
        if r >= R.shape[0]:
            break
        hit = np.nonzero(R[r:, c])[0]
        if hit.size == 0:
            continue
        k = r + hit[0]
        if k != r:
            R[[r, k]] = R[[k, r]]
        rows = np.nonzero(R[:, c])[0]
        rows = rows[rows != r]
        R[rows] ^= R[r]
        piv.append(c)
        r += 1
    return R[:r], piv


def gf2_nullspace(M):
    """Basis (rows) of {x : M x = 0 over GF(2)}."""
    M = (np.array(M, dtype=np.uint8) & 1)
    n = M.shape[1]
    R, piv = _gf2_rref(M)
    free = [c for c in range(n) if c not in set(piv)]
    N = np.zeros((len(free), n), dtype=np.uint8)
    for k, f in enumerate(free):
        N[k, f] = 1
        for r, pc in enumerate(piv):
            if R[r, f]:
                N[k, pc] = 1
    return N


def css_logicals(H_detect, H_other):
    """Logical operators that tell harmless residuals from logical errors for a CSS code whose errors are detected by
    H_detect: a basis of ker(H_other) modulo the row space of H_detect (k x n).  A residual r with H_detect r = 0 is a
    stabilizer iff L r = 0.  (Gross code: 12 rows; rotated surface code: 1.)"""
    Hd = np.asarray(sp.csc_matrix(H_detect).todense(), dtype=np.uint8) & 1
    Ho = np.asarray(sp.csc_matrix(H_other).todense(), dtype=np.uint8) & 1
    K = gf2_nullspace(Ho)                         # everything that commutes with the other check type
    S, _ = _gf2_rref(Hd)
    rank_s = S.shape[0]
    L = []
    cur = S.copy()
    for v in K:                                   # keep kernel vectors that enlarge the span of the stabilizers
        trial, _ = _gf2_rref(np.vstack([cur, v[None, :]]))
        if trial.shape[0] > cur.shape[0]:
            L.append(v)
            cur = trial
    assert cur.shape[0] == rank_s + len(L)
    return sp.csc_matrix(np.array(L, dtype=np.uint8).reshape(len(L), Hd.shape[1]))


def _shift(m):
    return sp.csc_matrix(np.roll(np.eye(m, dtype=np.int64), 1, axis=1))


def _bb_blocks(l=12, m=6):
    x = sp.kron(_shift(l), sp.identity(m, dtype=np.int64)).tocsc()
    y = sp.kron(sp.identity(l, dtype=np.int64), _shift(m)).tocsc()
    A = x ** 3 + y + y ** 2
    B = y ** 3 + x + x ** 2
    return A, B


def gross_x():
    """H_X = [A|B], A = x^3 + y + y^2, B = y^3 + x + x^2 on Z12 x Z6: 72 x 144, row weight 6."""
    A, B = _bb_blocks()
    H = sp.hstack([A, B]).tocoo()
    return _csc(H.row, H.col, H.shape)


def gross_z():
    """H_Z = [B^T|A^T]; only used to check H_X H_Z^T = 0."""
    A, B = _bb_blocks()
    H = sp.hstack([B.T, A.T]).tocoo()
    return _csc(H.row, H.col, H.shape)


def hgp_x(Hc):
    """H_X = [Hc (x) I_n | I_m (x) Hc^T] of the hypergraph product of Hc (m x n) with itself."""
    Hc = sp.csc_matrix(Hc).astype(np.int64)
    m, n = Hc.shape
    H = sp.hstack([sp.kron(Hc, sp.identity(n, dtype=np.int64)),
                   sp.kron(sp.identity(m, dtype=np.int64), Hc.T)]).tocoo()
    return _csc(H.row, H.col, H.shape)


SEED_H = 20240


def config_matrix(name):
    """The five BASELINE.json configs by short name (C1..C5) -> (H, per, max_iters)."""
    if name == "C1":
        return gallager(1000, 10, 9, seed=SEED_H + 1), 0.01, 25
    if name == "C2":
        return surface_x(15), 0.01, 32
    if name == "C3":
        return gross_x(), 0.01, 32
    if name == "C4":
        return hgp_x(gallager(32, 4, 3, seed=SEED_H + 4)), 0.02, 32
    if name == "C5":
        return gallager(100002, 6, 3, seed=SEED_H + 5), 0.02, 32
    raise KeyError(name)
