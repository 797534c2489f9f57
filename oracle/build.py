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


# ---------------------------------------------------------------------------------------------------
# oracle/_ref: the UNMODIFIED reference package, staged for the CPU arm of bench.py.
#
# minoring/NGU is pure Python over torch (no C/C++ to compile), so "building" the real reference here
# means staging its `ngu/` package next to a single-rank stand-in for mpi4py (absent from this image;
# `MPI.COMM_WORLD.Allreduce(send, recv)` == `recv[...] = send` is exactly one-rank MPI semantics for
# ngu/utils/mpi_util.py:17).  The copy is a BUILD OUTPUT: it lives only in the git-ignored oracle/_ref/
# (never in the repository's history), is made from the sources where they lie under /root/reference, and
# travels to the GPU box with the snapshot like the built .so files do.  Nothing under ngu_b200/ may
# import it (tests/test_cabi_cpu.py checks); bench.py uses it for `--impl reference` / `cpu_baseline`
# / the torch-eager GPU baseline only.
# ---------------------------------------------------------------------------------------------------
REF_SRC = os.environ.get("NGU_REFERENCE", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")

MPI_SHIM = '''"""single-rank stand-in for mpi4py (build output of oracle/build.py; not part of the reference)"""
class _Comm:
    def Allreduce(self, send, recv, op=None):
        recv[...] = send
    def Get_rank(self):
        return 0
    def Get_size(self):
        return 1
class MPI:
    SUM = "sum"
    COMM_WORLD = _Comm()
'''


def build_reference_copy(force=False):
    """Stage /root/reference/ngu -> oracle/_ref/ngu (+ the mpi4py stand-in).  Returns the directory, or None
    when the reference checkout is not present (the GPU box: it then uses what was staged here)."""
    marker = os.path.join(REF_DIR, "ngu", "models", "intrinsic_novelty", "intrinsic_novelty.py")
    src = os.path.join(REF_SRC, "ngu")
    if not os.path.isdir(src):
        return REF_DIR if os.path.exists(marker) else None
    if os.path.exists(marker) and not force:
        return REF_DIR
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    os.makedirs(REF_DIR)
    shutil.copytree(src, os.path.join(REF_DIR, "ngu"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.pt", "*.bin"))
    os.makedirs(os.path.join(REF_DIR, "mpi4py"))
    with open(os.path.join(REF_DIR, "mpi4py", "__init__.py"), "w") as f:
        f.write(MPI_SHIM)
    with open(os.path.join(REF_DIR, "README"), "w") as f:
        f.write("Build output of oracle/build.py: verbatim copy of the reference's ngu/ package (minoring/NGU) plus a\n"
                "single-rank mpi4py stand-in.  Git-ignored; used by bench.py's CPU arm only.\n")
    return REF_DIR


if __name__ == "__main__":
    print(build_reference_copy(force=True))
