#!/usr/bin/env python
"""Summarise an ncu report (raw page) and an ncu launch list into the text files kept under profiles/.
usage: tools/ncu_summary.py <report.ncu-rep> <out.txt> [launches.csv]"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:90s} {d[k]:>20s} {units[hdr.index(k)]}\n")
        if len(sys.argv) > 3:
            lr = list(csv.reader(open(sys.argv[3])))
            for i, r in enumerate(lr):
                if r and r[0] == "ID":
                    h, start = r, i + 1
                    break
            ik, iv = h.index("Kernel Name"), h.index("Metric Value")
            agg = collections.defaultdict(lambda: [0, 0.0])
            for r in lr[start:]:
                if len(r) <= iv:
                    continue
                try:
                    v = float(r[iv].replace(",", ""))
                except ValueError:
                    continue
                name = r[ik].split("(")[0][:70]
                agg[name][0] += 1
                agg[name][1] += v
            tot = sum(v[1] for v in agg.values())
            f.write(f"\n## launch list ({sys.argv[3].split('/')[-1]}): gpu__time_duration.sum per kernel, cold-cache serialised (compare SHARES)\n")
            for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
                f.write(f"{k:72s} n={v[0]:5d} total={v[1] / 1e6:10.3f} ms share={v[1] / tot * 100:5.1f}%\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
