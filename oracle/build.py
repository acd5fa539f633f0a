"""Builds the C part of the oracle (oracle/vq_oracle.c) with gcc.  Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libvq_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "vq_oracle.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB,
                           src, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
