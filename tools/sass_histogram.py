"""Per-kernel SASS opcode histogram of libracformer_ops.so (`cuobjdump -sass`): the evidence that the GEMM kernels are
tcgen05 / TMEM / TMA code (UTCHMMA, LDTM, UTMALDG, UBLKCP), that the gathers are 128-bit loads (LDG.E.128) and that the
scatter uses vector reductions (REDG.E.ADD.F32x4 ...). Runs without a GPU.

    python tools/sass_histogram.py > profiles/r02_sass_opcode_histogram.json
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "racformer_b200", "lib", "libracformer_ops.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDG.E.128",
         "LDG.E.64", "LDG.E", "STG.E.128", "STG.E", "LDS.128", "LDS", "STS.128", "STS", "LDGSTS", "REDG", "RED", "ATOMG",
         "FFMA2", "FFMA", "FMUL2", "FADD2", "HMMA", "SHFL", "BAR.SYNC", "MUFU", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = demangle(list(kernels))
    report = {}
    for k, ops in kernels.items():
        hist = collections.OrderedDict()
        for w in WATCH:
            n = sum(c for op, c in ops.items() if op == w or op.startswith(w + "."))
            if w in ("LDG.E", "STG.E", "LDS", "STS", "RED", "FFMA"):   # exclude the more specific rows counted above
                n = sum(c for op, c in ops.items() if (op == w or op.startswith(w + ".")) and not any(
                    op.startswith(x) for x in WATCH if x != w and x.startswith(w) and len(x) > len(w)))
            if n:
                hist[w] = n
        red = {op: c for op, c in ops.items() if op.startswith(("REDG", "RED.", "ATOMG"))}
        short = re.sub(r"\(.*", "", names.get(k, k))
        report[short] = {"instructions": sum(ops.values()), "opcodes": hist, **({"reductions": red} if red else {})}
    total = collections.Counter()
    for v in report.values():
        total.update(v["opcodes"])
    json.dump({"library": os.path.relpath(LIB, ROOT), "how": "cuobjdump -sass, opcode = first token of each instruction",
               "total": {k: total[k] for k in WATCH if total[k]}, "kernels": report}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
