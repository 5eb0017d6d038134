#!/usr/bin/env python
"""Builds a tuning variant of the CUDA core: tools/build_variant.py NAME -DHNM_X=1 ... -> _variants/NAME.so
(use with HNM_CORE_LIB=_variants/NAME.so; _variants/ is git-ignored but travels with gpurun)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "_variants"), exist_ok=True)
out = os.path.join(ROOT, "_variants", name + ".so")
cmd = [g.NVCC] + g.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-I", g.CSRC, os.path.join(g.CSRC, "hanamaru_b200.cu"), "-o", out]
subprocess.check_call(cmd)
print(out)
