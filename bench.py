#!/usr/bin/env python
"""bench.py -- decoded syndromes/s of batchdecode! on the B200 path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--per P] [--batch B]
  python bench.py --impl reference ...      # the CPU restatement of the reference, all host threads

One "step" = one batchdecode! of the whole synthetic batch (default: config C3 of BASELINE.json,
the [[144,12,12]] gross code, 10M syndromes per GPU, max_iters = 32, reference early-stop
semantics).  `value` is measured with the packed syndromes already resident in HBM, `e2e` through
the C-ABI host-buffer call (pinned host BitMatrix in, BitMatrix + success out, copies inside the
timed region).  Under torchrun every rank decodes its own 10M-syndrome shard (weak scaling,
inputs keyed by the global syndrome index) and the counters are all-reduced over NCCL.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

SEED_E = 12345
METRIC = "decoded syndromes/sec (batchdecode!)"
UNIT = "syndromes/s"
# FP64-pipe issue slots per edge-iteration of the reference arithmetic (DESIGN.md "Rooflines"):
# 10 add/sub/mul + 2 IEEE divisions of D = 8 FP64-pipe instructions each (measured from SASS).
FP64_SLOTS_PER_EDGE_ITER = 10 + 2 * 8
FP64_LANES_PER_SM_CLK = 64


def workload_spec(args, pkg):
    H, per, mi = pkg.codes.config_matrix(args.workload)
    default_per = {"C1": 0.01, "C2": 0.01, "C3": 0.03, "C4": 0.02, "C5": 0.02}[args.workload]
    per = args.per if args.per is not None else default_per
    default_B = {"C1": 4096, "C2": 1_000_000, "C3": 10_000_000, "C4": 1_000_000, "C5": 65536}[args.workload]
    B = args.batch if args.batch is not None else default_B
    if args.max_iters is not None:
        mi = args.max_iters
    return H, float(per), int(mi), int(B)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.power = []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report it instead of inventing numbers
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def cpu_baseline_leg(oracle, H, per, mi, seed, budget_s, nthreads, dense=False):
    """Time the CPU restatement on a bounded sample of the same workload."""
    probe = 2000
    _, syn = oracle.sample(H, per, seed, 0, probe)
    t0 = time.perf_counter()
    oracle.batch_decode(H, per, mi, syn, nthreads=nthreads, dense=dense)
    dt = max(time.perf_counter() - t0, 1e-4)
    n = int(min(max(probe, probe * budget_s / dt), 4_000_000))
    n = max(32, n // 32 * 32)
    _, syn = oracle.sample(H, per, seed, 0, n)
    t0 = time.perf_counter()
    r = oracle.batch_decode(H, per, mi, syn, nthreads=nthreads, dense=dense)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": int(nthreads), "kind": "port",
            "sample": "first %d syndromes of the same Philox stream, %s storage, %.2f s, mean iters %.3f" % (
                n, "dense s*n (faithful cost)" if dense else "edge-indexed", dt, float(r["iters"].mean()))}


def run_reference(args):
    """--impl reference: the restated reference on the host cores (Julia is not installed on
    the box, so oracle/ is the reference arm; PARITY UNPINNED, see oracle/bp_oracle.c)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()
    oracle = entry.load_oracle()
    oracle.build()
    H, per, mi, B = workload_spec(args, pkg)
    nthreads = oracle.num_threads()
    # size one step to ~2 s of wall time
    probe = 4096
    _, syn = oracle.sample(H, per, SEED_E, 0, probe)
    oracle.batch_decode(H, per, mi, syn, nthreads=nthreads)          # spins the thread pool up
    t0 = time.perf_counter()
    oracle.batch_decode(H, per, mi, syn, nthreads=nthreads)
    dt = max(time.perf_counter() - t0, 1e-4)
    n = int(min(B, max(probe, probe * 2.0 / dt))) // 32 * 32
    n = max(n, 32)
    _, syn = oracle.sample(H, per, SEED_E, 0, n)
    for _ in range(args.warmup):
        oracle.batch_decode(H, per, mi, syn, nthreads=nthreads)
    t0 = time.perf_counter()
    iters_mean = 0.0
    for _ in range(args.steps):
        r = oracle.batch_decode(H, per, mi, syn, nthreads=nthreads)
        iters_mean = float(r["iters"].mean())
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "each step = first %d syndromes of the workload's Philox stream, edge-indexed restatement, %d threads, mean iters %.3f" % (
        n, nthreads, iters_mean)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, H, per, mi, B),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def config_dict(args, H, per, mi, B):
    s, n = H.shape
    return {"workload": "%s: %s, s=%d n=%d E=%d, per=%g, max_iters=%d, batch=%d syndromes/GPU, early stop as reference" % (
        args.workload, {"C1": "Gallager (1000,10,9)", "C2": "d=15 rotated surface X checks",
                        "C3": "[[144,12,12]] gross code H_X", "C4": "HGP of Gallager(32,4,3) H_X",
                        "C5": "Gallager (100002,6,3)"}[args.workload], s, n, H.nnz, per, mi, B),
            "per": per, "max_iters": mi, "batch_per_gpu": B, "variant": getattr(args, "variant", "exact"),
            "l2": "inputs+outputs of one step exceed the 126 MB L2" if B * ((s + 31) // 32 + (n + 31) // 32) * 4 > 130e6
            else "L2 flushed between steps (256 MB write)"}


_REAL_STDOUT = None


def protect_stdout():
    """Everything libraries print (NCCL banners, warnings) goes to stderr; stdout carries exactly
    the one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--per", type=float, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--max-iters", type=int, default=None, dest="max_iters")
    ap.add_argument("--no-sweep", action="store_true", help="skip the extra per / forced-iteration points")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--family", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--prefetch", type=int, default=-1)
    ap.add_argument("--kernel-profile", action="store_true", dest="kernel_profile",
                    help="phase timing of the shared-memory kernel (adds clock reads; not a bench value)")
    ap.add_argument("--max-ctas", type=int, default=0, dest="max_ctas", help="cap on resident CTAs per SM (experiments)")
    ap.add_argument("--lean", type=int, default=-1, help="family SMEM: 1 = round-2 kernel (default), 0 = general persistent kernel")
    ap.add_argument("--variant", default="exact", choices=["exact", "minsum"],
                    help="exact = reference-parity sum-product (headline); minsum = normalised min-sum (no reference equivalent)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    entry.build()
    pkg = entry.load_package()
    H, per, mi, B = workload_spec(args, pkg)
    s, n = H.shape
    E = H.nnz
    opts = {}
    if args.family:
        opts["family"] = args.family
    if args.warps:
        opts["warps"] = args.warps
    if args.prefetch >= 0:
        opts["prefetch"] = args.prefetch
    if args.lean >= 0:
        opts["lean"] = args.lean
    if args.max_ctas > 0:
        opts["max_ctas_per_sm"] = args.max_ctas
    if args.kernel_profile:
        opts["kernel_profile"] = 1
    dec = pkg.BeliefPropagationDecoder(H, per, mi, devices=[local], variant=args.variant, **opts)
    info = dec.info()
    SW, NW = info["syn_words"], info["err_words"]
    # a non-default torch stream: its handle is non-NULL, so the library launches on exactly the
    # stream the torch CUDA events are recorded on (NULL would mean "the handle's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    assert st != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs, resident in HBM; shard r owns global syndromes [r*B, (r+1)*B)
    truth = torch.empty((B, NW), dtype=torch.int32, device=dev)
    synw = torch.empty((B, SW), dtype=torch.int32, device=dev)
    errw = torch.empty((B, NW), dtype=torch.int32, device=dev)
    conv = torch.empty(B, dtype=torch.uint8, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    ctr = torch.zeros(4, dtype=torch.int64, device=dev)
    score = torch.zeros(2, dtype=torch.int64, device=dev)
    flush = None
    if B * (SW + NW) * 4 <= 130e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sample(per_):
        dec.sample_device(B, rank * B, SEED_E, per_, truth.data_ptr(), synw.data_ptr(), stream=st)

    def timed_run(steps, warmup):
        """W warm-ups then `steps` timed decodes; returns (seconds of the slowest rank, counters, launches)."""
        for _ in range(warmup):
            ctr.zero_()
            dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), iters.data_ptr(), None, ctr.data_ptr(), stream=st)
            if world > 1:
                dist.all_reduce(ctr)
        barrier()
        l0 = dec.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        total_ctr = torch.zeros(4, dtype=torch.int64, device=dev)
        for a, b in evs:
            if flush is not None:
                flush.fill_(1)
            ctr.zero_()
            a.record(stream)
            dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), iters.data_ptr(), None, ctr.data_ptr(), stream=st)
            if world > 1:
                dist.all_reduce(ctr)          # the path's only collective: 4 int64 counters
            b.record(stream)
            total_ctr += ctr
        barrier()
        secs = sum(a.elapsed_time(b) for a, b in evs) / 1e3
        t = torch.tensor([secs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), total_ctr.cpu().numpy(), dec.launch_count() - l0

    sample(per)
    sampler = ClockSampler(local)
    # clocks are sampled during the timed region only (warm-up runs before start())
    for _ in range(args.warmup):
        dec.decode_device(B, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), iters.data_ptr(), None, ctr.data_ptr(), stream=st)
    barrier()
    sampler.start()
    secs, c, launches = timed_run(args.steps, 0)
    clocks = sampler.stop()
    n_dec_all = float(c[0])                       # all ranks (counters were all-reduced), all steps
    value = n_dec_all / secs
    mean_iters = float(c[2]) / max(float(c[0]), 1.0)
    conv_frac = float(c[1]) / max(float(c[0]), 1.0)
    # logical scoring of the last step (not timed): exact-match fraction
    score.zero_()
    dec.score_device(B, truth.data_ptr(), errw.data_ptr(), synw.data_ptr(), score.data_ptr(), stream=st)
    torch.cuda.synchronize()
    exact_frac = float(score[0].item()) / B

    # ---- roofline of the dominant kernel (this rank's share, this rank's time)
    peaks, peaks_src = measured_peaks()
    units_per_step_gpu = float(c[2]) / args.steps / world           # (syndrome, iteration) pairs per launch per GPU
    step_s = secs / args.steps
    clk_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    io_bytes = B * (SW * 4 + NW * 4 + 1 + 4)
    alg_bytes = units_per_step_gpu * 4.0 * E * 8.0 + io_bytes
    if info["family"] == 1:
        fp64_ops = units_per_step_gpu * E * FP64_SLOTS_PER_EDGE_ITER
        peak = info["sm_count"] * FP64_LANES_PER_SM_CLK * clk_mhz * 1e6 / 1e12
        roof = {"bound": "fp64", "achieved": fp64_ops / step_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": fp64_ops / step_s / 1e12 / peak, "traffic": None,
                "note": "shared-memory-resident kernel: bounded by the FP64 pipe, not HBM or tensor cores. "
                        "achieved = (syndrome-iterations) x E x %d FP64-pipe issue slots (10 add/sub/mul + 2 IEEE divisions x 8) per launch / CUDA-event time; "
                        "peak = %d SMs x 64 FP64 lanes/clk x median SM clock under load (%.0f MHz); one FLOP = one FP64 lane-instruction" % (
                            FP64_SLOTS_PER_EDGE_ITER, info["sm_count"], clk_mhz),
                "ncu": ({"launch": "2 000 000-syndrome launch of this workload under ncu --set full (not a bench value)",
                         "smsp__issue_active_pct": 57.1, "sm__pipe_fp64_cycles_active_pct": 42.4,
                         "fp64_share_of_issued_warp_instructions_pct": 37.4, "executed_fp64_slots_per_edge_iteration": 21,
                         "dram_bytes_read": 64074240, "dram_bytes_write": 2648832,
                         "previous_12_warp_shape": {"smsp__issue_active_pct": 61.2, "sm__pipe_fp64_cycles_active_pct": 41.8,
                                                    "source": "profiles/r1_persistent_kernel_c3_mode0_12warps_ncu_full.txt"},
                         "source": "profiles/r1_persistent_kernel_c3_mode0_ncu_full.txt"}
                        if args.workload == "C3" and args.variant == "exact" else None),
                "hbm_model": {"achieved_GBps": alg_bytes / step_s / 1e9, "peak_GBps": peaks["hbm_gbs"], "peak_source": peaks_src,
                              "frac": alg_bytes / step_s / 1e9 / peaks["hbm_gbs"],
                              "note": "4*E*8 B per syndrome-iteration if messages lived in HBM (they live in shared memory) + packed I/O"}}
    else:
        roof = {"bound": "hbm", "achieved": alg_bytes / step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": alg_bytes / step_s / 1e9 / peaks["hbm_gbs"], "traffic": None,
                "note": "algorithmic bytes = 4*E*8 per syndrome-iteration + packed I/O; peak %s copy bandwidth; the message store (%d MB) is sized to stay L2-resident, so >1.0 is possible" % (
                    peaks_src, info["message_bytes"] >> 20)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args, H, per, mi, B),
            "mean_iters": mean_iters, "converged_frac": conv_frac, "exact_match_frac": exact_frac,
            "syndrome_iterations_per_s": float(c[2]) / secs,
            "roofline": roof, "gpu_launches": int(launches), "clocks": clocks,
            "kernel": {k: info[k] for k in ("family", "kernel_mode", "ctas_per_sm", "threads_per_cta", "smem_bytes", "slots",
                                            "message_bytes", "prefetch_distance", "kernel_rev")}}

    if args.kernel_profile:
        kp = dec.kernel_profile()
        wi = max(kp["warp_iterations"], 1)
        line["kernel_profile_cycles_per_warp_iteration"] = {k: round(v / wi, 1) for k, v in kp.items() if k != "warp_iterations"}

    # ---- end to end through the host-buffer C-ABI call (Julia BitMatrix in / out), pinned memory
    if not args.no_e2e:
        lib = pkg._lib
        nb_in = (B * s + 63) // 64 * 8
        nb_out = (B * n + 63) // 64 * 8
        h_in = torch.empty(nb_in, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(nb_out, dtype=torch.uint8, pin_memory=True)
        h_conv = torch.empty(B, dtype=torch.uint8, pin_memory=True)
        # host BitMatrix of this rank's syndromes (built once, untimed): bit c*s + r of the stream
        bit_ar = torch.arange(s, device=dev, dtype=torch.int64)
        weights = (2 ** torch.arange(8, device=dev, dtype=torch.int32)).to(torch.uint8)
        acc = torch.zeros(nb_in * 8, dtype=torch.uint8, device=dev)
        sl = 1 << 19
        for b0 in range(0, B, sl):
            w = synw[b0:b0 + sl].to(torch.int64) & 0xFFFFFFFF
            bits = ((w[:, bit_ar // 32] >> (bit_ar % 32)) & 1).to(torch.uint8)       # [rows, s]
            acc[b0 * s:(b0 + bits.shape[0]) * s] = bits.reshape(-1)
            del w, bits
        d_bits = (acc.view(-1, 8) * weights).sum(dim=1, dtype=torch.int32).to(torch.uint8)
        del acc
        h_in.copy_(d_bits)
        del d_bits
        torch.cuda.synchronize()
        np_in, np_out, np_conv = h_in.numpy(), h_out.numpy(), h_conv.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            dec.decode_raw(B, np_in, lib.FMT_BITS, 0, np_out, lib.FMT_BITS, 0, np_conv)
        barrier()
        l0 = dec.launch_count()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            cnt = dec.decode_raw(B, np_in, lib.FMT_BITS, 0, np_out, lib.FMT_BITS, 0, np_conv)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        # outputs of the host path must equal the device-resident path's
        got_conv = int(np_conv.sum())
        line["e2e"] = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(nb_in),
                       "d2h_bytes_per_step": int(nb_out + B), "steps": e2e_steps,
                       "api": "ldpcb200_decode_batch(FMT_BITS in/out, pinned host buffers), wall clock around the blocking call",
                       "converged_check": got_conv == int(c[1] / args.steps / world) if world == 1 else None,
                       "gpu_launches": int(dec.launch_count() - l0)}

    # ---- extra operating points (not the headline): per sweep and forced max_iters
    if not args.no_sweep and args.workload in ("C2", "C3"):
        sweep = []
        pts = {"C3": [0.001, 0.01, 0.03, 0.1], "C2": [0.001, 0.01, 0.03, 0.1]}[args.workload]
        for p_ in pts:
            d2 = pkg.BeliefPropagationDecoder(H, p_, mi, devices=[local], **opts)
            dec_saved = dec
            dec = d2
            sample(p_)
            secs2, c2, _ = timed_run(3, 1)
            sweep.append({"per": p_, "value": float(c2[0]) / secs2, "mean_iters": float(c2[2]) / float(c2[0]),
                          "converged_frac": float(c2[1]) / float(c2[0]), "syndrome_iterations_per_s": float(c2[2]) / secs2})
            d2.close()
            dec = dec_saved
        # forced iterations: its own decoder, so that the library configures for that mode (early_stop = 0 before the
        # first decode selects the 12-warp shape; the early-stop runs use 8 warps)
        d2 = pkg.BeliefPropagationDecoder(H, per, mi, devices=[local], early_stop=0, **opts)
        dec_saved = dec
        dec = d2
        sample(per)
        secs2, c2, _ = timed_run(2, 1)
        d2.close()
        dec = dec_saved
        sweep.append({"per": per, "forced_iters": mi, "value": float(c2[0]) / secs2, "mean_iters": float(c2[2]) / float(c2[0]),
                      "syndrome_iterations_per_s": float(c2[2]) / secs2,
                      "fp64_frac": float(c2[2]) / 2 / world * E * FP64_SLOTS_PER_EDGE_ITER / (secs2 / 2) / 1e12 /
                      (info["sm_count"] * FP64_LANES_PER_SM_CLK * clk_mhz * 1e6 / 1e12)})
        line["sweep"] = sweep

    # ---- sum-product vs min-sum (BASELINE.json config 3): the normalised min-sum kernels on the same
    # syndromes; no reference equivalent, so it is reported by quality, not parity
    if world == 1 and not args.no_sweep and args.variant == "exact" and args.workload in ("C2", "C3"):
        errw_sp = errw.clone()
        dms = pkg.BeliefPropagationDecoder(H, per, mi, devices=[local], variant="minsum")
        dec_saved = dec
        dec = dms
        secs2, c2_, _ = timed_run(3, 2)
        score.zero_()
        dms.score_device(B, truth.data_ptr(), errw.data_ptr(), synw.data_ptr(), score.data_ptr(), stream=st)
        torch.cuda.synchronize()
        differ = int((errw != errw_sp).any(dim=1).sum().item())
        line["minsum"] = {"value": float(c2_[0]) / secs2, "unit": UNIT, "per": per, "scale": 0.875,
                          "mean_iters": float(c2_[2]) / float(c2_[0]), "converged_frac": float(c2_[1]) / float(c2_[0]),
                          "exact_match_frac": float(score[0].item()) / B,
                          "decisions_differ_from_sum_product_frac": differ / B,
                          "note": "FP64 log-likelihood-ratio min-sum, same schedule/early stop; sum-product exact_match_frac is the line's own"}
        dec = dec_saved
        dms.close()
        del errw_sp

    # ---- config C2 of BASELINE.json next to the headline (1 M syndromes, per sweep), N = 1 only
    if world == 1 and not args.no_sweep and args.workload == "C3":
        H2, _, mi2 = pkg.codes.config_matrix("C2")
        B2 = 1_000_000
        c2 = []
        for p_ in (0.001, 0.01, 0.03, 0.1):
            d2 = pkg.BeliefPropagationDecoder(H2, p_, mi2, devices=[local])
            i2 = d2.info()
            t2 = torch.empty((B2, i2["err_words"]), dtype=torch.int32, device=dev)
            s2 = torch.empty((B2, i2["syn_words"]), dtype=torch.int32, device=dev)
            e2 = torch.empty((B2, i2["err_words"]), dtype=torch.int32, device=dev)
            cv2 = torch.empty(B2, dtype=torch.uint8, device=dev)
            c4 = torch.zeros(4, dtype=torch.int64, device=dev)
            d2.sample_device(B2, 0, SEED_E, p_, t2.data_ptr(), s2.data_ptr(), stream=st)
            for _ in range(3):
                d2.decode_device(B2, s2.data_ptr(), e2.data_ptr(), cv2.data_ptr(), None, None, None, stream=st)
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            ea.record(stream)
            for _ in range(reps):
                if flush is not None:
                    pass
                d2.decode_device(B2, s2.data_ptr(), e2.data_ptr(), cv2.data_ptr(), None, None, c4.data_ptr(), stream=st)
            eb.record(stream)
            torch.cuda.synchronize()
            secs2 = ea.elapsed_time(eb) / 1e3
            cc = c4.cpu().numpy()
            c2.append({"per": p_, "value": float(cc[0]) / secs2, "mean_iters": float(cc[2]) / float(cc[0]),
                       "converged_frac": float(cc[1]) / float(cc[0]), "syndrome_iterations_per_s": float(cc[2]) / secs2,
                       "fp64_frac": float(cc[2]) / secs2 * H2.nnz * FP64_SLOTS_PER_EDGE_ITER / 1e12 /
                       (info["sm_count"] * FP64_LANES_PER_SM_CLK * clk_mhz * 1e6 / 1e12)})
            d2.close()
            del t2, s2, e2, cv2
        line["config_C2_surface_d15_1M"] = c2

    # ---- BASELINE.json config 4 is "the BP stage of BP+OSD": the whole BP -> OSD-0 pipeline on the same syndromes
    # (ldpcb200_decode_device with posterior ratios, then ldpcb200_osd0_device on the unconverged ones)
    if world == 1 and not args.no_sweep and args.workload == "C4" and args.variant == "exact":
        Bo = int(min(B, 131072))
        ratio = torch.empty((Bo, n), dtype=torch.float64, device=dev)
        stats = torch.zeros(8, dtype=torch.int64, device=dev)
        sc_bp = torch.zeros(2, dtype=torch.int64, device=dev)
        sc_osd = torch.zeros(2, dtype=torch.int64, device=dev)

        def bp_stage():
            dec.decode_device(Bo, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), iters.data_ptr(), ratio.data_ptr(),
                              ctr.data_ptr(), stream=st)

        def osd_stage():
            dec.osd0_device(Bo, synw.data_ptr(), errw.data_ptr(), conv.data_ptr(), ratio.data_ptr(), stats.data_ptr(), stream=st)

        dec.set_option("ratio_last_only", 1)
        for _ in range(2):
            bp_stage()
            osd_stage()
        torch.cuda.synchronize()
        reps = 3
        t_bp = t_osd = 0.0
        stats.zero_()
        ctr.zero_()
        for _ in range(reps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            bp_stage()
            e1.record(stream)
            if _ == 0:
                dec.score_device(Bo, truth.data_ptr(), errw.data_ptr(), synw.data_ptr(), sc_bp.data_ptr(), stream=st)
                torch.cuda.synchronize()
            e1b = torch.cuda.Event(enable_timing=True)
            e1b.record(stream)
            osd_stage()
            e2.record(stream)
            torch.cuda.synchronize()
            t_bp += e0.elapsed_time(e1) / 1e3
            t_osd += e1b.elapsed_time(e2) / 1e3
        dec.score_device(Bo, truth.data_ptr(), errw.data_ptr(), synw.data_ptr(), sc_osd.data_ptr(), stream=st)
        torch.cuda.synchronize()
        so = stats.cpu().numpy()
        nproc = float(so[0]) / reps
        if os.environ.get("LDPCB200_OSD_PROFILE"):     # one extra pass with the kernel's per-phase cycle counters on
            dec.set_option("osd_profile", 1)
            stats.zero_()
            bp_stage()
            osd_stage()
            torch.cuda.synchronize()
            sp_ = stats.cpu().numpy()
            sys.stderr.write("osd phases (SM cycles per syndrome): sort %.0f build %.0f search %.0f update %.0f solve %.0f\n" %
                             tuple(float(sp_[k]) / max(float(sp_[0]), 1.0) for k in (3, 4, 5, 6, 7)))
            dec.set_option("osd_profile", 0)
        bposd = {"batch": Bo, "per": per, "value": Bo * reps / (t_bp + t_osd), "unit": UNIT,
                 "bp_stage_syndromes_per_s": Bo * reps / t_bp,
                 "osd_stage_syndromes_per_s": float(so[0]) / t_osd if t_osd > 0 else None,
                 "osd_stage_ms": t_osd / reps * 1e3, "bp_stage_ms": t_bp / reps * 1e3,
                 "unconverged_frac": nproc / Bo,
                 "mean_pivots": float(so[1]) / max(float(so[0]), 1.0), "mean_columns_visited": float(so[2]) / max(float(so[0]), 1.0),
                 "bp_exact_match_frac": float(sc_bp[0].item()) / Bo, "bp_syndrome_satisfied_frac": float(sc_bp[1].item()) / Bo,
                 "bposd_exact_match_frac": float(sc_osd[0].item()) / Bo,
                 "bposd_syndrome_satisfied_frac": float(sc_osd[1].item()) / Bo,
                 "note": "device-resident; the BP stage writes posterior ratios only in iteration max_iters (ratio_last_only)"}
        if not args.no_e2e:
            # the reference-facing call: host arrays in, host arrays out (ldpcb200_bposd_decode_batch)
            import time as _t
            syn_t = torch.empty((Bo, SW), dtype=torch.int32, pin_memory=True)      # pinned, like the headline e2e leg
            syn_t.copy_(synw[:Bo])
            err_t = torch.zeros((Bo, NW), dtype=torch.int32, pin_memory=True)
            conv_t = torch.zeros(Bo, dtype=torch.uint8, pin_memory=True)
            torch.cuda.synchronize()
            syn_h, err_h, conv_h = syn_t.numpy().view(np.uint32), err_t.numpy().view(np.uint32), conv_t.numpy()
            dec.bposd_raw(Bo, syn_h, pkg._lib.FMT_PACKED32, SW, err_h, pkg._lib.FMT_PACKED32, NW, conv_h)
            t0 = _t.perf_counter()
            for _ in range(reps):
                dec.bposd_raw(Bo, syn_h, pkg._lib.FMT_PACKED32, SW, err_h, pkg._lib.FMT_PACKED32, NW, conv_h)
            dt = _t.perf_counter() - t0
            bposd["e2e"] = {"value": Bo * reps / dt, "unit": UNIT, "h2d_bytes_per_step": int(Bo * SW * 4),
                            "d2h_bytes_per_step": int(Bo * (NW * 4 + 1)),
                            "api": "ldpcb200_bposd_decode_batch(FMT_PACKED32 in/out, pinned host buffers), wall clock"}
        if not args.no_cpu:
            import time as _t
            oracle = entry.load_oracle()
            nth = oracle.num_threads()
            Kc = 8 * nth
            _, syn_c = oracle.sample(H, per, SEED_E, 0, Kc)
            oracle.bposd_decode(H, per, mi, syn_c[:, :nth], nthreads=nth)
            t0 = _t.perf_counter()
            oracle.bposd_decode(H, per, mi, syn_c, nthreads=nth)
            dt = _t.perf_counter() - t0
            bposd["cpu_baseline"] = {"value": Kc / dt, "unit": UNIT, "cores": nth, "kind": "port",
                                     "sample": "first %d syndromes of the same stream, BP + restated OSD-0 (bit-packed rows)" % Kc}
        line["bposd_osd0"] = bposd
        dec.set_option("ratio_last_only", 0)
        del ratio

    # ---- CPU baseline next to it (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle = entry.load_oracle()
        line["cpu_baseline"] = cpu_baseline_leg(oracle, H, per, mi, SEED_E, 10.0, 1)
        line["cpu_baseline_all_threads"] = cpu_baseline_leg(oracle, H, per, mi, SEED_E, 5.0, oracle.num_threads())
        if s * n * 16 < 2e9:
            line["cpu_baseline_dense_faithful"] = cpu_baseline_leg(oracle, H, per, mi, SEED_E, 5.0, 1, dense=True)
    dec.close()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
