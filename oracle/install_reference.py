#!/usr/bin/env python
"""Recipe for the reference arm of bench.py: copies the UNMODIFIED reference tree (pure Python, ~1 MB) from
/root/reference into baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot,
where /root/reference does not exist).  Run by __graft_entry__.build() whenever /root/reference is present.
Nothing is edited; the shims the 2021 code needs on this image live in oracle/ref_loader.py, outside the copy.

    python oracle/install_reference.py [--src /root/reference] [--dst baseline/_ref]
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(src="/root/reference", dst=os.path.join(ROOT, "baseline", "_ref")):
    if not os.path.isdir(os.path.join(src, "anomaly")):
        return None
    for sub in ("anomaly", "DeepLabV3Plus-Pytorch"):
        out = os.path.join(dst, sub)
        if os.path.isdir(out):
            shutil.rmtree(out)
        shutil.copytree(os.path.join(src, sub), out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
    for f in ("README.md", "requirements.txt", "LICENSE"):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    print(install(a.src, a.dst))
