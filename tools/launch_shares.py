"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import json
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, total, n = {}, 0.0, 0
    for r in rows[h + 1 + skip:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "")[:80]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
        total += v
        n += 1
    top = sorted(agg.items(), key=lambda kv: -kv[1][0])
    ours = {k: v for k, v in agg.items() if "racf::" in k or k.startswith("racf")}
    print(json.dumps({"launches": n, "total_us": total,
                      "ours_share": sum(v[0] for v in ours.values()) / total if total else None,
                      "ours": {k: {"us": v[0], "calls": v[1], "share": v[0] / total} for k, v in ours.items()},
                      "top": [{"name": k, "us": v[0], "calls": v[1], "share": v[0] / total} for k, v in top[:15]]}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
