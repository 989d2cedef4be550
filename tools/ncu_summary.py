"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep files stay in gpurun_out/).

    python tools/ncu_summary.py rep gpurun_out/prof_conv_tc.ncu-rep profiles/r01_conv_tc_full.md
    python tools/ncu_summary.py launches gpurun_out/launches_simt_step.csv profiles/r01_launches_simt.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
]


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu --set full summary of %s" % path, ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("## launch %s: %s  grid %s block %s" % (r[hdr.index("ID")], name[:90],
                                                            r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?",
                                                            r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
        for k in KEYS:
            if k in hdr:
                lines.append("- %s = %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


def launches(path, out, top=40):
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    hdr = rows[0]
    ni, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ui = hdr.index("Metric Unit")
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1000.0
        elif r[ui] in ("ms", "msecond"):
            v *= 1000.0
        name = r[ni].split("(")[0][:80]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    lines = ["# kernel launch list summary of %s" % path,
             "", "total launches %d, total device time %.3f ms (serialised, cold-cache: compare SHARES)" %
             (sum(a[0] for a in agg.values()), total / 1000.0), "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append("| %s | %d | %.1f | %.3f |" % (name, n, us, us / total))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
