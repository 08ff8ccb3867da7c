"""Build the CPU oracle (test infrastructure) into oracle/_build/liboracle.so.

    python oracle/build.py

gcc only; -ffp-contract=off so the only fused multiply-adds are the fmaf() calls the
restatement spells out (they mirror nvcc's contraction of the reference kernels).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pointnet2_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fopenmp", "-fvisibility=hidden",
           "-shared", "-fPIC", SRC, "-o", OUT, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
