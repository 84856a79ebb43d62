#!/usr/bin/env python
"""Tuning aid: builds libntedit_b200.so with extra compiler flags into gpurun_variants/<name>/ (travels to the GPU box, stays
out of git); load it with NTB_LIB=gpurun_variants/<name>/libntedit_b200.so.
usage: python tools/build_variant.py <name> [-DNTB_X=1 ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntedit_b200 import lib  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(ROOT, "gpurun_variants", name)
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libntedit_b200.so")
    cmd = [os.environ.get("NVCC", "nvcc")] + lib.NVCC_FLAGS + flags + ["-Xcompiler", "-fPIC", "-shared"] + lib._sources() + ["-o", so]
    subprocess.run(cmd, check=True)
    print(so)


if __name__ == "__main__":
    main()
