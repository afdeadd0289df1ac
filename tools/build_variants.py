#!/usr/bin/env python
"""Builds tuning variants of libhesaff_b200.so with extra -D flags into hesaff_b200/variants/ (git-ignored *.so; they
travel to the GPU box).  Select one at run time with HESAFF_LIB=hesaff_b200/variants/<name>.so.
   python tools/build_variants.py name1="-DA=1 -DB=2" name2="..." """
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hesaff_b200 import build as B
out = os.path.join(ROOT, "hesaff_b200", "variants")
os.makedirs(out, exist_ok=True)
procs = []
for a in sys.argv[1:]:
    name, flags = a.split("=", 1)
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + B.NVCC_FLAGS + flags.split() + \
        ["-o", os.path.join(out, name + ".so")] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    print(name, "rc", p.wait())
