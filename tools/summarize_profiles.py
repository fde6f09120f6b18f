#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.
    python tools/summarize_profiles.py <tag> <prof.ncu-rep> <launches.csv> "<command profiled>"
Runs here (no GPU): ncu -i ... --page raw / --page source."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
           "sm__cycles_active.avg",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def ncu(*args):
    return subprocess.run(["ncu"] + list(args), capture_output=True, text=True).stdout


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * mult.get(unit, 1)


def full_set(tag, rep, cmd):
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = data[0][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "k_search_step"
    out = ["# %s `%s` -- `ncu --set full --clock-control none --import-source on`" % (tag, name), "",
           "Command profiled: `%s`." % cmd,
           "ncu flushes caches between replays and serialises kernels: these are cold-L2 numbers.", "",
           "| metric | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |",
           "|---|" + "---|" * len(data)]
    traffic = []
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append("| `%s` | " % m + " | ".join("%s %s" % (r[i], units[i]) for r in data) + " |")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in data:
        traffic.append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
    mean_traffic = sum(traffic) / len(traffic)
    out += ["", "DRAM traffic per launch (read+write), mean of %d launches: **%.2f MB**." % (len(traffic), mean_traffic / 1e6), ""]
    # source page: stall samples and executed instructions per source line
    src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                                          "--launch-skip", "0", "--launch-count", "1"))))
    cur, h, agg = None, None, []
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if r[0] == "Line No":
            h = r; continue
        if r[0] in ("Function Name",) or h is None or r[0] == "":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(h[4:], r[4:]))
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
        agg.append((int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0), cur, ln, r[1].strip()[:100], stalls))
    tot_s, tot_i = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    by_reason = defaultdict(int)
    for a in agg:
        for k, v in a[5].items():
            by_reason[k] += v
    out += ["Warp-stall samples by reason (first launch, %d samples, %d warp-instructions executed): " % (tot_s, tot_i) +
            ", ".join("%s %.0f%%" % (k, 100.0 * v / tot_s) for k, v in sorted(by_reason.items(), key=lambda x: -x[1])[:8]) + ".", "",
            "Top source lines by stall samples:", "", "| samples | instr | line | source | top stalls |", "|---|---|---|---|---|"]
    for s_, i_, f, ln, text, st in sorted(agg, reverse=True)[:16]:
        top = ", ".join("%s %d" % kv for kv in sorted(st.items(), key=lambda x: -x[1])[:2])
        out.append("| %.1f%% | %.2f%% | %s:%d | `%s` | %s |" % (100.0 * s_ / tot_s, 100.0 * i_ / tot_i, f, ln, text.replace("|", "\\|"), top))
    open(os.path.join(ROOT, "profiles", "%s_k_search_step_full.md" % tag), "w").write("\n".join(out) + "\n")
    return mean_traffic


def launch_list(tag, path, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = defaultdict(list)
    for r in rows:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            u = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
            per[r["Kernel Name"]].append(v)
    tot = sum(sum(v) for v in per.values())
    n = sum(len(v) for v in per.values())
    out = ["# %s launch list" % tag, "", "Command: `%s`" % cmd,
           "(%d consecutive launches inside the CUDA-graph replays of one UCT_search; per-launch times are cold-cache and "
           "serialised: compare shares.)" % n, "", "| share | launches | avg us | kernel |", "|---|---|---|---|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append("| %.1f%% | %d | %.1f | `%s` |" % (100.0 * sum(v) / tot, len(v), sum(v) / len(v), k[:110]))
    mine = sum(sum(v) for k, v in per.items() if "dbaz::" in k)
    step = sum(sum(v) for k, v in per.items() if "k_search_step" in k)
    out += ["", "Total %.0f us over %d launches.  `k_search_step` = %.1f%% of the time, all engine kernels (`dbaz::*`) = %.1f%%, "
            "library kernels (cuDNN / cuBLASLt) = %.1f%%." % (tot, n, 100 * step / tot, 100 * mine / tot, 100 * (tot - mine) / tot)]
    open(os.path.join(ROOT, "profiles", "%s_launches_summary.md" % tag), "w").write("\n".join(out) + "\n")
    return step / tot


if __name__ == "__main__":
    tag, rep, launches, cmd = sys.argv[1:5]
    traffic = full_set(tag, rep, cmd)
    share = launch_list(tag, launches, cmd)
    json.dump({"k_search_step_dram_bytes_per_launch": traffic, "k_search_step_share_of_launch_list": share,
               "source": "profiles/%s_k_search_step_full.md (ncu --set full, mean of the captured launches)" % tag},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))
    print("traffic %.2f MB/launch, k_search_step share of the launch list %.1f%%" % (traffic / 1e6, 100 * share))
