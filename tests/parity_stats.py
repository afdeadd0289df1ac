#!/usr/bin/env python
"""GPU box helper: prints compare_keypoints() of the CUDA path against the plain-C oracle port on a few textures
(identical-descriptor fraction, worst deviations).  python tests/parity_stats.py [WxH:seed ...]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hesaff_b200
from oracle import oracle
from tools.gen_textured import textured
from parity import compare_keypoints

cases = [a for a in sys.argv[1:]] or ["640x480:1", "400x300:33"]
port = oracle.load("port")
for c in cases:
    wh, seed = c.split(":")
    w, h = (int(v) for v in wh.split("x"))
    img = textured(w, h, int(seed))
    det = hesaff_b200.AffineHessianDetector(hesaff_b200.HessianAffineParams(), device=0, max_width=w, max_height=h, max_batch=1)
    det.detectPyramidKeypoints(img)
    got = det.keys()
    want = port.detect(img.astype(np.float32))
    st = compare_keypoints(got, want[want["described"] == 1])
    print(c, "det", int(det.n_detected[0]), len(want), json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in st.items()}), flush=True)
    det.close()
