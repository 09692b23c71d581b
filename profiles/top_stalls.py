"""Top instructions by stall samples from an `ncu --page source --csv` dump.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > src.csv; python profiles/top_stalls.py src.csv [n]"""
import csv
import sys


def main(path, n=25):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or not r[0].startswith("0x"):
            if body:
                break                      # next kernel instance in the same dump
            continue
        body.append(r)
    col = {h: i for i, h in enumerate(hdr)}
    samp = col["# Samples"]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[samp] or 0) for r in body)
    print("total samples", total, "instructions", len(body), "executed(sum)", sum(int(r[col["Instructions Executed"]] or 0) for r in body))
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall mix:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:n]
    for i in sorted(order):
        r = body[i]
        st = sorted(((int(r[j] or 0), hdr[j]) for j in stall_cols), reverse=True)[:2]
        print("%5d  %6s  x%-8s %-70s %s" % (i, r[samp], r[col["Instructions Executed"]], r[col["Source"]].strip()[:70], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
