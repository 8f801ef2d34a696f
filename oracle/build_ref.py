"""Recipe for oracle/_ref: the UNMODIFIED reference sources, copied where they can travel — TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (run by __graft_entry__.build() whenever /root/reference is present)

The reference is pure Python (tasks/R2R-judy/src, no build system), so "building" it is copying its package next to
the oracle: oracle/_ref/tasks/R2R-judy/src/**.py.  oracle/_ref/ is git-ignored (reference sources never enter this
repository's history) but not gpurun-ignored, so it reaches the GPU box like a built .so does.  There
oracle/ref_loader.py falls back to it, and `bench.py --impl reference` / `cpu_baseline` time the reference's OWN
EnvDropAgent.rollout + trainer iteration (kind "reference") instead of the oracle port.  Nothing in the product
package imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("VLN_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
REL = os.path.join("tasks", "R2R-judy", "src")


def build(verbose=True):
    src = os.path.join(SRC_ROOT, REL)
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/_ref: {src} not present, nothing to do")
        return None
    dst = os.path.join(DST_ROOT, REL)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for root, _dirs, files in os.walk(src):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            out_dir = os.path.join(dst, os.path.relpath(root, src))
            os.makedirs(out_dir, exist_ok=True)
            shutil.copyfile(os.path.join(root, fn), os.path.join(out_dir, fn))
            n += 1
    with open(os.path.join(DST_ROOT, "README"), "w") as f:
        f.write("Unmodified copy of the reference's tasks/R2R-judy/src (made by oracle/build_ref.py); git-ignored.\n")
    if verbose:
        print(f"oracle/_ref: copied {n} reference source files to {dst}")
    return dst


if __name__ == "__main__":
    sys.exit(0 if build() or True else 1)
