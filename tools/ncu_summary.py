#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the handful of numbers DESIGN.md / profiles cite."""
import csv, json, os, subprocess, sys, io
# usage: ncu_summary.py report.ncu-rep [sass_listing.txt] [--json out.json key=value ...]
#   --json writes the figures bench.py quotes in `roofline.ncu` (with the commit and whatever key=value pairs describe
#   the profiled launch: workload, variant, syndromes ...), so that the bench line never carries hand-copied numbers.
json_out, json_meta = None, {}
if "--json" in sys.argv:
    k = sys.argv.index("--json")
    json_out = sys.argv[k + 1]
    for kv in sys.argv[k + 2:]:
        a, b = kv.split("=", 1)
        json_meta[a] = b
    sys.argv = sys.argv[:k]
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
summary = {}
for row in rows:
    print("=" * 100)
    for k in keys + stall:
        if k in hdr:
            i = hdr.index(k)
            if not summary.get("_done"):
                summary[k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")] = (row[i], units[i])
            print("%-90s %s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), row[i], units[i]))
    summary["_done"] = True
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
    summary["_mix"] = {"total_warp_instructions": tot, "fp64_pipe_warp_instructions": fp64,
                       "top": {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]}}
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
        # stall-reason breakdown of the most-sampled instructions (every per-instruction sampling column ncu exports)
        reason_cols = [c for c in h if c.startswith("stall_") or "Stall" in c and c != col]
        with open(sys.argv[2] + ".reasons", "w") as f:
            f.write("columns: %s\n" % ", ".join(h))
            for k, sm, src_ in sorted(lines, key=lambda l: -l[1])[:60]:
                x = body[k]
                parts = []
                for c in h:
                    if c in ("Address", "Source", "Instructions Executed", col) or c not in ix:
                        continue
                    v = x[ix[c]]
                    try:
                        fv = float(v)
                    except ValueError:
                        continue
                    if fv != 0 and (c.startswith("stall") or "stall" in c.lower()):
                        parts.append("%s=%s" % (c, v))
                f.write("#%d %d %s | %s\n" % (k, sm, src_.strip(), " ".join(parts)))
        print("-" * 100)
        print("top instructions by %s:" % col)
        for k, sm, src_ in sorted(lines, key=lambda l: -l[1])[:30]:
            print("  #%-5d %7d %5.1f%%  %s" % (k, sm, 100.0 * sm / tots, src_))

if json_out:
    def num(k):
        v = summary.get(k)
        if not v:
            return None
        try:
            x = float(v[0].replace(",", ""))
        except ValueError:
            return None
        u = v[1].lower()
        scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
                 "nsecond": 1e-9, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        return x * scale
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
    except Exception:
        commit = ""
    out = {"source": "ncu --set full --clock-control none, one launch (not a bench value)", "commit": commit or json_meta.get("commit", ""),
           "kernel": summary.get("Kernel Name", ("", ""))[0].strip(),
           "duration_s": num("gpu__time_duration.sum"),
           "registers_per_thread": num("launch__registers_per_thread"),
           "sm__pipe_fp64_cycles_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "smsp__issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "sm__warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
           "sm__inst_executed_pipe_lsu_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
           "sm__inst_executed_pipe_xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
           "l1tex__throughput_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
           "dram__throughput_pct": num("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
           "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
           "smsp__inst_executed": num("smsp__inst_executed.sum"),
           "stalls_per_issue": {k[6:]: float(v[0]) for k, v in summary.items() if k.startswith("stall:") and v[0] not in ("", "n/a")},
           "instruction_mix": summary.get("_mix")}
    mix = summary.get("_mix")
    if mix and mix["total_warp_instructions"]:
        out["fp64_share_of_issued_warp_instructions_pct"] = 100.0 * mix["fp64_pipe_warp_instructions"] / mix["total_warp_instructions"]
    out.update(json_meta)
    with open(json_out, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
