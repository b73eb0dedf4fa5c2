"""Compile oracle/racf_oracle.c -> oracle/libracf_oracle.so with gcc (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "racf_oracle.c")
LIB = os.path.join(HERE, "libracf_oracle.so")
CFLAGS = ["-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-Wall"]


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # some images wrap gcc in a way that loses libgomp.spec: try the plain system compiler too, and as a
    # last resort build single-threaded
    log = ""
    for flags in (CFLAGS, [f for f in CFLAGS if f != "-fopenmp"]):
        for cc in ("/usr/bin/gcc", "gcc", os.environ.get("CC") or "cc"):
            cmd = [cc] + flags + ["-o", LIB, SRC, "-lm"]
            try:
                res = subprocess.run(cmd, capture_output=True, text=True)
            except OSError as e:
                log += f"{cc}: {e}\n"
                continue
            if res.returncode == 0:
                return LIB
            log += " ".join(cmd) + "\n" + res.stdout + res.stderr
    raise RuntimeError("could not compile the oracle:\n" + log)


if __name__ == "__main__":
    print(build(force=True))
