#!/usr/bin/env python
"""ncu report (.ncu-rep from `ncu --set full --clock-control none --import-source on`) -> a tracked markdown summary.
    python tools/ncu_md.py <report.ncu-rep> <profiles/out.md> "<title>" "<command profiled>" [notes...]
Runs here (no GPU): ncu -i ... --page raw / --page source."""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_tmem_ldt.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def ncu(*args):
    return subprocess.run(["ncu"] + list(args), capture_output=True, text=True).stdout


def main():
    rep, out_path, title, cmd = sys.argv[1:5]
    notes = sys.argv[5:]
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kname = data[0][hdr.index("Kernel Name")]
    out = ["# %s" % title, "", "Kernel `%s`; command profiled: `%s` under `ncu --set full --clock-control none --import-source on`." % (kname, cmd),
           "ncu serialises kernels, replays each one and flushes caches between replays: durations are cold-cache figures at the profiler's",
           "clock, never bench values.", "", "| metric | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |", "|---|" + "---|" * len(data)]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append("| `%s` | " % m + " | ".join(("%s %s" % (r[i], units[i])).strip() for r in data) + " |")
    # source page: the lines with the most stall samples
    src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--launch-skip", "0", "--launch-count", "1"))))
    h = None
    lines = []
    for r in src:
        if "Source" in r and any("Sampling" in c for c in r):
            h = r
            continue
        if h and len(r) == len(h):
            try:
                s_col = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
                samples = float(r[s_col[0]]) if s_col else 0.0
            except ValueError:
                continue
            lines.append((samples, r[h.index("Source")].strip(), r[h.index("#")] if "#" in h else ""))
    if lines:
        tot = sum(s for s, _, _ in lines) or 1.0
        out += ["", "Source lines with the most warp-stall samples (share of all samples):", ""]
        for s, text, ln in sorted(lines, reverse=True)[:14]:
            if s <= 0:
                break
            out.append("* %4.1f %%  line %s: `%s`" % (100.0 * s / tot, ln, text[:150]))
    if notes:
        out += [""] + notes
    open(out_path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
