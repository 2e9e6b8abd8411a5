#!/usr/bin/env python
"""Per-source-line roll-up of an `ncu --set full --import-source on` capture (scratch tool).

ncu's CSV source page is per SASS instruction without line numbers; nvdisasm -g prints the same
instructions of the same cubin with their file/line.  The two are joined by instruction order.

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_ID(0-based launch in the report) SO_FILE KERNEL_SUBSTRING [top]
"""
import collections
import csv
import io
import pathlib
import re
import subprocess
import sys
import tempfile


def sass_lines(so: str, kernel_sub: str):
    tmp = pathlib.Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(pathlib.Path(so).resolve())], cwd=tmp, check=True,
                   capture_output=True)
    for cubin in sorted(tmp.glob("*.cubin")):
        out = subprocess.run(["nvdisasm", "-g", str(cubin)], capture_output=True, text=True).stdout
        cur, loc, res, found = None, ("?", 0), [], False
        for line in out.splitlines():
            m = re.match(r"\.text\.(\S+):", line)
            if m:
                if found:
                    return res
                cur = m.group(1)
                found = kernel_sub in cur and "correspond" in cur or (kernel_sub in cur)
                loc = ("?", 0)
                continue
            if not found:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                loc = (pathlib.Path(m.group(1)).name, int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
                res.append(loc)
        if found:
            return res
    raise SystemExit("kernel not found in " + so)


def main():
    rep, kid, so, ksub = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    # the page holds one table per launch, each introduced by a "Kernel Name" row
    tables, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = []
            tables.append(cur)
            continue
        if cur is not None:
            cur.append(row)
    tab = tables[kid]
    hdr, rows = tab[0], tab[1:]
    col = {h: i for i, h in enumerate(hdr)}
    locs = sass_lines(so, ksub)
    if len(locs) != len(rows):
        print(f"warning: {len(rows)} profiled instructions vs {len(locs)} disassembled", file=sys.stderr)
    agg = collections.defaultdict(lambda: collections.Counter())
    keys = ["Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_long_sb", "stall_wait",
            "stall_short_sb", "stall_branch_resolving", "stall_not_selected", "stall_barrier", "stall_lg"]
    for loc, r in zip(locs, rows):
        for k in keys:
            if k in col:
                try:
                    agg[loc][k] += float(r[col[k]])
                except ValueError:
                    pass
    tot = collections.Counter()
    for c in agg.values():
        tot.update(c)
    print("totals:", {k: int(v) for k, v in tot.items()})
    src_cache = {}

    def src(loc):
        f, ln = loc
        if f not in src_cache:
            cands = list(pathlib.Path(__file__).resolve().parent.parent.rglob(f))
            src_cache[f] = cands[0].read_text().splitlines() if cands else []
        L = src_cache[f]
        return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
    print(f"{'file:line':28s} {'inst%':>6s} {'thr/inst':>8s} {'smp%':>6s} {'long_sb%':>8s}  source")
    for loc, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        ie = c["Instructions Executed"]
        print(f"{loc[0] + ':' + str(loc[1]):28s} {100 * ie / max(1, tot['Instructions Executed']):6.2f} "
              f"{c['Thread Instructions Executed'] / max(1, ie):8.1f} {100 * c['# Samples'] / max(1, tot['# Samples']):6.2f} "
              f"{100 * c['stall_long_sb'] / max(1, tot['# Samples']):8.2f}  {src(loc)}")


if __name__ == "__main__":
    main()
