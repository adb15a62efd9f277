"""
Recipe for oracle/_ref: a byte-for-byte copy of the UNMODIFIED reference's Python sources
(/root/reference -> oracle/_ref/ppo_and_friends), so that the reference itself can run on the GPU
box (which has no /root/reference) as the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind
= "reference"`).  oracle/_ref/ is git-ignored (it is an output, like a compiled .so) but NOT
gpurun-ignored, so it travels with the snapshot.  Nothing under ppo_and_friends_b200/ imports it.

    python oracle/build_ref.py            # no-op when /root/reference is absent (the GPU box)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PPOAF_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "ppo_and_friends")


def build(verbose=False):
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__", "test", "docs", "images")]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if not f.endswith(".py"):
                continue
            out_dir = os.path.join(DST, rel) if rel != "." else DST
            os.makedirs(out_dir, exist_ok=True)
            dst = os.path.join(out_dir, f)
            src = os.path.join(root, f)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or \
                    os.path.getsize(dst) != os.path.getsize(src):
                shutil.copyfile(src, dst)
            n += 1
    if verbose:
        print(f"oracle/_ref: {n} reference source files under {DST}")
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
