"""Compile oracle/tb200_oracle.c with gcc into oracle/liboracle.so (test infra)."""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tb200_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


def build_oracle(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found; cannot build the C oracle")
    # -ffp-contract=off: no fused multiply-adds the source does not spell out
    subprocess.run(
        [gcc, "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
         "-o", LIB, SRC, "-lm"],
        check=True,
    )
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
