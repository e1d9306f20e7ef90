#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the handful of numbers DESIGN.md / profiles cite."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = r[0], r[1], r[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for row in rows:
    print("=" * 100)
    for k in keys + stall:
        if k in hdr:
            i = hdr.index(k)
            print("%-90s %s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), row[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
if his:
    h = rows[his[0]]
    body = rows[his[0] + 1: his[1] - 1 if len(his) > 1 else None]
    ix = {n: i for i, n in enumerate(h)}
    agg, tot = {}, 0
    for x in body:
        if len(x) < len(h):
            continue
        try:
            ins = int(x[ix["Instructions Executed"]])
        except ValueError:
            continue
        toks = x[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        agg[op] = agg.get(op, 0) + ins
        tot += ins
    print("-" * 100)
    print("dynamic warp-instruction mix of the first launch (total %d):" % tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:24]:
        print("  %-10s %14d %5.1f%%" % (k, v, 100.0 * v / tot))
    fp64 = sum(v for k, v in agg.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print("  FP64-pipe share of issued instructions: %.1f%%" % (100.0 * fp64 / tot))
    # per-instruction stall samples (SASS listing with sample counts; top lines first)
    col = next((c for c in ("Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)", "# Samples") if c in ix), None)
    if col and len(sys.argv) > 2:
        lines = []
        for k, x in enumerate(body):
            if len(x) < len(h):
                continue
            try:
                sm = int(x[ix[col]])
            except ValueError:
                sm = 0
            try:
                ex = int(x[ix["Instructions Executed"]])
            except ValueError:
                ex = 0
            lines.append((k, sm, "%10d  %s" % (ex, x[ix["Source"]])))
        tots = sum(l[1] for l in lines) or 1
        with open(sys.argv[2], "w") as f:
            f.write("# SASS listing: index, %s, share, warp-instructions executed, instruction (total samples %d)\n" % (col, tots))
            for k, sm, src_ in lines:
                f.write("%5d %7d %5.1f%%  %s\n" % (k, sm, 100.0 * sm / tots, src_))
        print("-" * 100)
        print("top instructions by %s:" % col)
        for k, sm, src_ in sorted(lines, key=lambda l: -l[1])[:30]:
            print("  #%-5d %7d %5.1f%%  %s" % (k, sm, 100.0 * sm / tots, src_))
