#!/usr/bin/env python
"""Stages the UNMODIFIED Python reference for timing on the GPU box (TEST / BENCH INFRASTRUCTURE).

The reference is pure Python: there is nothing to compile into oracle/_ref, and /root/reference does not exist on the
GPU box.  This recipe copies - byte for byte, from where they lie under /root/reference - the package
(src/hydromodel/*.py) and the gerd_roseires case (its modules and the CSV tables it reads) into oracle/_ref/pyref/,
which is git-ignored (nothing of it enters the history) but travels with the repo snapshot.  bench.py's cpu_baseline leg
then runs a few members of the calibration ensemble through the reference's own PreissmannSolver.run on the box's host
cores (oracle/ref_harness.py, PR_REFERENCE_ROOT) - "the reference CPU path ... in the same run" of the north star.

    python oracle/stage_reference.py          (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

SRC = os.environ.get("PR_REFERENCE_SRC", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "pyref")

KEEP = (".py", ".csv")


def stage() -> bool:
    if not os.path.isdir(os.path.join(SRC, "src", "hydromodel")):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for sub in ("src", os.path.join("cases", "gerd_roseires")):
        for root, dirs, files in os.walk(os.path.join(SRC, sub)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "raw")]
            for f in files:
                if not f.endswith(KEEP):
                    continue
                rel = os.path.relpath(os.path.join(root, f), SRC)
                out = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), out)
                n += 1
    init = os.path.join(SRC, "cases", "__init__.py")
    if os.path.exists(init):
        shutil.copyfile(init, os.path.join(DST, "cases", "__init__.py"))
    print(f"staged {n} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
