"""
Build recipe for libppoaf_b200.so (hand-written sm_100a CUDA behind a C ABI, include/ppoaf_b200.h).

    python -m ppo_and_friends_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is built IN-TREE next to the sources so it travels with
the repo snapshot to the GPU box (it is git-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(CSRC, "libppoaf_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")

SOURCES = ["segscan.cu", "gather.cu", "stats.cu", "mlp.cu", "loss.cu", "optim.cu", "step.cu", "peer.cu", "fused_step.cu"]
HEADERS = ["common.cuh", "mlp.cuh", "umma.cuh", "loss_common.cuh", "internal.h", os.path.join(ROOT, "include", "ppoaf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
    "-Xlinker", "-rpath,/usr/local/cuda/lib64",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libppoaf_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
