"""Compile the plain-C oracle (oracle/episodic_oracle.c) with gcc -> oracle/_build/liboracle_epi.so.

TEST INFRASTRUCTURE ONLY.  -ffp-contract=off keeps the separately rounded
multiply/add the restatement needs; no -ffast-math.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "episodic_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_epi.so")


def build_oracle(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found; the C oracle cannot be built")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [gcc, "-O3", "-mavx2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
