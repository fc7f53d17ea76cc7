#!/usr/bin/env python
"""Stages a VERBATIM copy of the reference's Python package and tests under baseline/_ref/ (git-ignored, travels to
the GPU box with the tree) so that the drop-in acceptance tests (tests/test_dropin_reference.py) can run the reference's
own, unmodified test files and examples with deodr_b200 bound as `deodr.differentiable_renderer_cython`.
Nothing under baseline/_ref/ is product source; the Cython extension of the reference is NOT built there."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def stage() -> bool:
    if not os.path.isdir(os.path.join(SRC, "deodr")):
        return False
    os.makedirs(DST, exist_ok=True)
    for sub in ("deodr", "tests"):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so"))
    for dirpath, _, files in os.walk(DST):
        os.chmod(dirpath, 0o755)
        for f in files:
            os.chmod(os.path.join(dirpath, f), 0o644)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here", DST)
    sys.exit(0)
