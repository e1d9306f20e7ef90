#!/usr/bin/env python
"""One BP -> OSD-0 pass on config C4 (device-resident), small batch: the command ncu wraps to capture osd0_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
H, per, mi = pkg.codes.config_matrix("C4")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dec = pkg.BeliefPropagationDecoder(H, per, mi, devices=[0])
info = dec.info()
dev = torch.device("cuda:0")
SW, NW, n = info["syn_words"], info["err_words"], H.shape[1]
truth = torch.empty((B, NW), dtype=torch.int32, device=dev)
syn = torch.empty((B, SW), dtype=torch.int32, device=dev)
err = torch.empty((B, NW), dtype=torch.int32, device=dev)
conv = torch.empty(B, dtype=torch.uint8, device=dev)
ratio = torch.empty((B, n), dtype=torch.float64, device=dev)
stats = torch.zeros(8, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
dec.set_option("ratio_last_only", 1)
dec.sample_device(B, 0, 12345, per, truth.data_ptr(), syn.data_ptr())
dec.decode_device(B, syn.data_ptr(), err.data_ptr(), conv.data_ptr(), None, ratio.data_ptr(), None)
dec.osd0_device(B, syn.data_ptr(), err.data_ptr(), conv.data_ptr(), ratio.data_ptr(), stats.data_ptr())
torch.cuda.synchronize()
print("osd stats", stats.cpu().numpy())
dec.close()
