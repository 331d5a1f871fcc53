#!/usr/bin/env python
"""Copy the reference's Python packages needed to run ITS model zoo into git-ignored baseline/_ref/ so that
tools/bench_reference_gpu.py can run the UNMODIFIED reference on the GPU box (where /root/reference does not exist).
SURVEY.md section 8c sanctions exactly this ("copy the needed files into git-ignored baseline/_ref/ before the call, never
commit them").  Nothing under baseline/_ref/ is tracked, imported by the product, or used by the -m gpu tests."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("COTB200_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(os.path.join(REF, "cupy_layers")):
        print("make_ref_snapshot: no reference tree at %s; leaving %s as it is" % (REF, DST))
        return 0
    os.makedirs(DST, exist_ok=True)
    for pkg in ("cupy_layers", "models", "config", "utils", "optim"):
        dst = os.path.join(DST, pkg)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    print("make_ref_snapshot: %s" % DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
