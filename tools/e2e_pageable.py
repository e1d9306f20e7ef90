"""End-to-end ldpcb200_decode_batch on C3 with BitMatrix buffers in pinned vs ordinary (pageable) host memory, with and
without overlapping chunk kernels: wall clock around the blocking call."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
pkg = entry.load_package(); oracle = entry.load_oracle(); lib = pkg._lib
H, _, mi = pkg.codes.config_matrix("C3")
per = float(sys.argv[1]) if len(sys.argv) > 1 else 0.03
s, n = H.shape
B = 10_000_000
rng = np.random.default_rng(1)
nb_in = (B * s + 63) // 64 * 8
nb_out = (B * n + 63) // 64 * 8
# sparse random syndromes (weight like per 0.03) are not needed for timing the host path: use a real sample of 1M, tiled
_, syn = oracle.sample(H, per, 3, 0, 1_000_000)
bits = np.packbits(np.asfortranarray(syn).T.reshape(-1), bitorder="little")
src = np.tile(bits, 10)[:nb_in].copy()
for overlap, direct in ((1, 1), (2, 1), (0, 1), (1, 0), (2, 0), (0, 0)):
    for chunks in (0,):
        opts = dict(overlap_chunks=overlap, direct_bits=direct)
        if chunks:
            opts["chunk"] = (B // chunks + 31) // 32 * 32
        dec = pkg.BeliefPropagationDecoder(H, per, mi, **opts)
        for mem in ("pinned",) if chunks else ("pinned", "pageable"):
            if mem == "pinned":
                h_in = torch.empty(nb_in, dtype=torch.uint8, pin_memory=True); h_in.numpy()[:] = src
                h_out = torch.empty(nb_out, dtype=torch.uint8, pin_memory=True)
                h_cv = torch.empty(B, dtype=torch.uint8, pin_memory=True)
                a_in, a_out, a_cv = h_in.numpy(), h_out.numpy(), h_cv.numpy()
            else:
                a_in, a_out, a_cv = src.copy(), np.zeros(nb_out, np.uint8), np.zeros(B, np.uint8)
            for _ in range(2):
                dec.decode_raw(B, a_in, lib.FMT_BITS, 0, a_out, lib.FMT_BITS, 0, a_cv)
            t0 = time.perf_counter()
            for _ in range(5):
                dec.decode_raw(B, a_in, lib.FMT_BITS, 0, a_out, lib.FMT_BITS, 0, a_cv)
            dt = (time.perf_counter() - t0) / 5
            print("per %g direct_bits=%d overlap_chunks=%d chunks=%s %-8s %.3e syndromes/s (%.2f ms per 10 M)" % (per, direct, overlap, chunks or "auto", mem, B / dt, dt * 1e3))
        dec.close()
