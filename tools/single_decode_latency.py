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

# kernel-level view: forced 25 iterations (early stop off), device-resident, B = 1 and B = 100, both kernels
import torch
dev = torch.device("cuda:0")
info = dec.info()
SW, NW = info["syn_words"], info["err_words"]
for B in (1, 100):
    truth = torch.empty((B, NW), dtype=torch.int32, device=dev)
    synw = torch.empty((B, SW), dtype=torch.int32, device=dev)
    errw = torch.empty((B, NW), dtype=torch.int32, device=dev)
    conv = torch.empty(B, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    dec.sample_device(B, 0, 2024, per, truth.data_ptr(), synw.data_ptr(), stream=st)
    out = []
    for early in (1, 0):
        dec.set_option("early_stop", early)
        for sb in (0, -1):
            dec.set_option("small_batch", sb)
            for _ in range(5):
                dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), None, None, None, stream=st)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(50):
                dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), None, None, None, stream=st)
            b_.record(stream)
            torch.cuda.synchronize()
            out.append("%s/%s %.1f us" % ("early-stop" if early else "forced-25", "persistent" if sb == 0 else "node-parallel",
                                          a.elapsed_time(b_) / 50 * 1e3))
    print("C1 device-resident decode of %d syndrome(s): %s" % (B, "; ".join(out)))
dec.close()

# ---- the same question for a LARGE code (C5, n = 100 002: the messages of one syndrome, 2.4 MB, do not fit in an SM's
# shared memory): grid-wide cooperative kernel vs one lane of the persistent kernel, device-resident, early stop
H, per, mi = pkg.codes.config_matrix("C5")
dec = pkg.BeliefPropagationDecoder(H, per, mi)
info = dec.info()
SW, NW = info["syn_words"], info["err_words"]
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
st = stream.cuda_stream
for B in (1, 16):
    truth = torch.empty((B, NW), dtype=torch.int32, device=dev)
    synw = torch.empty((B, SW), dtype=torch.int32, device=dev)
    errw = torch.empty((B, NW), dtype=torch.int32, device=dev)
    conv = torch.empty(B, dtype=torch.uint8, device=dev)
    its = torch.empty(B, dtype=torch.int32, device=dev)
    dec.sample_device(B, 0, 2024, per, truth.data_ptr(), synw.data_ptr(), stream=st)
    out = []
    for gk, reps in ((1, 20), (0, 3)):
        dec.set_option("grid_kernel", gk)
        for _ in range(2):
            dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), its.data_ptr(), None, None, stream=st)
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), its.data_ptr(), None, None, stream=st)
        b_.record(stream)
        torch.cuda.synchronize()
        out.append("%s %.1f us (mean iterations %.2f)" % ("grid-wide kernel" if gk else "persistent kernel", a.elapsed_time(b_) / reps * 1e3,
                                                          float(its.float().mean().item())))
    print("C5 device-resident decode of %d syndrome(s): %s" % (B, "; ".join(out)))
syn1 = oracle.sample(H, per, 2024, 0, 4)[1]
for b in range(2):
    pkg.decode_b(dec.__class__(H, per, mi) if False else dec, syn1[:, b])
dec.set_option("grid_kernel", 1)
t0 = time.perf_counter()
for b in range(4):
    pkg.decode_b(dec, syn1[:, b])
print("C5 single decode! through the Python mirror (host vectors in/out, posterior ratios back): %.1f us per call" % ((time.perf_counter() - t0) / 4 * 1e6))
dec.close()
