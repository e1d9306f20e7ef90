"""decode!(decoder, syndrome) latency on config C1 (Gallager (1000,10,9), per 0.01, 25 iterations):
GPU path through the Python mirror vs the restated reference (dense 'faithful cost' and edge-indexed)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); oracle = entry.load_oracle()
H, per, mi = pkg.codes.config_matrix("C1")
errs, syn = oracle.sample(H, per, 2024, 0, 200)
dec = pkg.BeliefPropagationDecoder(H, per, mi)
for b in range(20):
    pkg.decode_b(dec, syn[:, b])
t0 = time.perf_counter()
for b in range(200):
    g, ok = pkg.decode_b(dec, syn[:, b])
gpu = (time.perf_counter() - t0) / 200
t0 = time.perf_counter(); oracle.batch_decode(H, per, mi, syn, dense=True); dense = (time.perf_counter() - t0) / 200
t0 = time.perf_counter(); oracle.batch_decode(H, per, mi, syn); edge = (time.perf_counter() - t0) / 200
print("C1 single decode!: GPU (ctypes mirror, host vectors in/out) %.1f us; CPU restatement dense %.1f us, edge-indexed %.1f us" % (gpu * 1e6, dense * 1e6, edge * 1e6))
