"""Build recipe for the oracle's C restatement (TEST INFRASTRUCTURE ONLY).

Compiles oracle/prox_oracle.c into oracle/_build/libprox_oracle.so with gcc.
Nothing from /root/reference is compiled: the reference is pure Python
(SURVEY.md section 2a), so there is no oracle/_ref binary for this repo -- the
"reference run here" leg is oracle/ref_harness.py, which imports the reference
modules in place (this container only).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libprox_oracle.so")
SRC = os.path.join(HERE, "prox_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
