"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [steps]"""
import collections
import csv
import sys


def main(path, steps=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("nsecond", "ns"):
            v /= 1000.0
        elif row["Metric Unit"] in ("msecond", "ms"):
            v *= 1000.0
        agg.setdefault(row["Kernel Name"].split("(")[0][:56], []).append(v)
    ours = {k: v for k, v in agg.items() if "ipsr::" in k or "innercos" in k}
    per_step = {k: (sum(v) / len(v)) for k, v in ours.items() if len(v) >= 3}
    total = sum(per_step.values())
    print("| kernel | launches | mean us | min us | max us | share of step |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, v in agg.items():
        share = "%.1f %%" % (100 * per_step[k] / total) if k in per_step else "-"
        print("| `%s` | %d | %.2f | %.2f | %.2f | %s |" % (k, len(v), sum(v) / len(v), min(v), max(v), share))
    print("\nsum of per-step kernel means: %.1f us" % total)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
