"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share of the step.  usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = name.replace("<unnamed>::", "")
        name = re.sub(r"<.*", "", name)
        name = re.sub(r"\(.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    out = ["| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k in sorted(tot, key=tot.get, reverse=True):
        out.append(f"| {k[:80]} | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / total:.1f}% | {tot[k] / cnt[k]:.2f} |")
    out.append(f"| **total** | {sum(cnt.values())} | {total:.1f} | 100% | |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
