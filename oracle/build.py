"""Build the oracle's C restatement into oracle/_build/libnavsim_oracle.so.

TEST INFRASTRUCTURE ONLY (see the header of navsim_oracle.c).  IEEE-strict
flags: -O2 -ffp-contract=off, no -ffast-math, so double arithmetic matches the
reference's Cython build (setuptools default -O2, x86-64 baseline, no FMA).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "navsim_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libnavsim_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-std=c11", "-D_GNU_SOURCE", "-Wall", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
