"""Summarise an .ncu-rep (read here, on the CPU box): key raw metrics, stall reasons, hottest source lines.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "?")[:100])
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active"]
    for k in keys:
        if k in d:
            print(f"  {k:72s} {d[k]:>16s} {u[k]}")
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
for i, r in enumerate(rows[:12]):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
if hdr:
    iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    lines, tot, totS = [], 0, 0
    for r in rows[start:]:
        if len(r) > iI and r[0] != "" and r[2] == "-":
            try:
                n, s = int(r[iI]), int(r[iS])
            except ValueError:
                continue
            lines.append((n, s, r[0], r[1]))
            tot += n
            totS += s
    print(f"source lines by stall samples (total inst {tot}, samples {totS}):")
    for n, s, ln, code in sorted(lines, key=lambda x: -x[1])[:nl]:
        print(f"  {s / max(totS, 1) * 100:5.1f}% smp {n / max(tot, 1) * 100:5.1f}% inst  L{ln}: {code.strip()[:105]}")
