#!/usr/bin/env python
"""Turns the scratch artefacts of one tools/gpu_check.sh run (gpurun_out/<tag>_*) into the tracked
summaries under profiles/ (round prefix given on the command line).

    python tools/make_profiles.py <tag> <round prefix, e.g. r01>

  <prefix>_bench_n1.json                     the N=1 bench line
  <prefix>_launches_bench_steps2.csv         ncu launch list (gpu__time_duration.sum, --clock-control none)
  <prefix>_launches_summary.md               per-kernel totals / shares of that list
  <prefix>_correspond_ncu_full_summary.csv   selected `ncu --set full` metrics per captured launch
  traffic.json                               DRAM bytes per launch of the correspondence kernel
"""
import collections
import csv
import io
import json
import pathlib
import shutil
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_config_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = name.split("(")[0]
    return name if len(name) < 70 else name[:67] + "..."


def main():
    tag, prefix = sys.argv[1], sys.argv[2]
    src = ROOT / "gpurun_out"
    OUT.mkdir(exist_ok=True)
    bench = (src / f"{tag}_bench.json").read_text().strip().splitlines()[-1]
    json.loads(bench)
    (OUT / f"{prefix}_bench_n1.json").write_text(bench + "\n")

    launches = src / f"{tag}_launches.csv"
    if launches.exists():
        shutil.copy(launches, OUT / f"{prefix}_launches_bench_steps2.csv")
        rows = [r for r in csv.reader(launches.open()) if len(r) > 5]
        hdr = next(r for r in rows if "Kernel Name" in r)
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows[rows.index(hdr) + 1:]:
            try:
                v = float(r[vi].replace(",", "")) / 1000.0  # ns -> us
            except ValueError:
                continue
            agg[short(r[ki])][0] += 1
            agg[short(r[ki])][1] += v
        total = sum(v[1] for v in agg.values())
        lines = [f"source: gpurun_out/{launches.name} ({sum(v[0] for v in agg.values())} launches, {total:.1f} us of "
                 "kernel time; cold-cache, serialised - compare shares, not absolutes)", "",
                 "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f} % |")
        (OUT / f"{prefix}_launches_summary.md").write_text("\n".join(lines) + "\n")

    rep = src / f"{tag}_prof_corr.ncu-rep"
    if rep.exists():
        txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        cols = [hdr.index(m) for m in METRICS if m in hdr]
        names = [hdr.index("Kernel Name")] + cols
        with (OUT / f"{prefix}_correspond_ncu_full_summary.csv").open("w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] for i in names])
            w.writerow([units[i] for i in names])
            for r in data:
                w.writerow([r[i] for i in names])
        ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        real = [r for r in data if float(r[ti]) > 20.0]  # speculative launches past convergence return at once
        per = [float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]] for r in real]
        (OUT / "traffic.json").write_text(json.dumps({
            "correspond_kernel_dram_bytes_per_launch": int(sum(per) / len(per)),
            "source": f"profiles/{prefix}_correspond_ncu_full_summary.csv: mean over the {len(per)} real launches of "
                      "dram__bytes_read.sum + dram__bytes_write.sum from one ncu --set full --clock-control none capture "
                      f"(gpurun_out/{rep.name}, scratch; ncu flushes the caches before every replay, so this is the "
                      "cold-L2 traffic - in a match the target tree stays L2-resident across iterations)",
            "algorithmic_bytes_per_launch": 40000000}, indent=1) + "\n")
    print("profiles refreshed from tag", tag)


if __name__ == "__main__":
    main()
