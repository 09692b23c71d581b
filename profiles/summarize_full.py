"""Summarise an `ncu --set full` report exported with `ncu -i rep.ncu-rep --page raw --csv > raw.csv`.
usage: python profiles/summarize_full.py raw.csv [out.json]
Prints a markdown table (one row per captured launch) and, optionally, writes per-kernel DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch) as JSON for bench.py's roofline.traffic."""
import csv
import json
import sys

COLS = [
    ("gpu__time_duration.sum", "time us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor mem %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
]


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main(path, out_json=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    present = [(k, n) for k, n in COLS if k in col]
    tensor_cols = [h for h in hdr if "tensor" in h and "pct_of_peak" in h]
    print("| kernel | " + " | ".join(n for _, n in present) + " |")
    print("|---|" + "---:|" * len(present))
    traffic = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:48]
        vals = []
        for k, _ in present:
            v = to_float(r[col[k]])
            u = units[col[k]]
            if v is not None and u == "byte":
                v /= 1e6
            elif v is not None and u == "Kbyte":
                v /= 1e3
            elif v is not None and u == "Gbyte":
                v *= 1e3
            elif v is not None and u in ("nsecond", "ns"):
                v /= 1e3
            elif v is not None and u in ("msecond", "ms"):
                v *= 1e3
            vals.append("-" if v is None else ("%.2f" % v if abs(v) < 1000 else "%.0f" % v))
        print("| `%s` | " % name + " | ".join(vals) + " |")
        if "dram__bytes_read.sum" in col:
            def mb(k):
                v, u = to_float(r[col[k]]), units[col[k]]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
                return (v or 0.0) * scale
            traffic.setdefault(name, []).append(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
    if tensor_cols:
        print("\ntensor-pipe metrics present:", ", ".join(tensor_cols[:6]))
    if out_json:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
