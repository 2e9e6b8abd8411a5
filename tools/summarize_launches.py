"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    head, rows = rows[0], rows[1:]
    ik, iv = head.index("Kernel Name"), head.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[ik])
        name = name.replace("<unnamed>::", "")
        name = re.sub(r"<.*", "<...>", name)[:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path} ({len(rows)} launches, {tot / 1e3:.1f} us of kernel time; cold-cache, serialised)\n")
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / tot:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])
