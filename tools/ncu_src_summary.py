"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest instructions."""
import csv
import sys


def main(path, thresh=0.012):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    idx = {name: i for i, name in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):   # the export repeats per launch / per view: keep the first
            break
        if len(r) == len(hdr):
            data.append(r)
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    print("total samples", tot, "instructions", len(data))
    reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[idx[n]] or 0) for r in data) for n in reasons}
    for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print(f"{n:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
    print()
    for i, r in enumerate(data):
        s = int(r[idx["# Samples"]] or 0)
        if s > tot * thresh:
            top = max(reasons, key=lambda n: int(r[idx[n]] or 0))
            print(f"{i:4d} {s:6d} {100 * s / tot:5.1f}% {top:18s} {r[idx['Source']].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.012)
