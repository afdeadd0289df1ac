#!/usr/bin/env python
"""Generate tests/golden/ from the REFERENCE build (oracle/_ref: perdoch/hesaff sources compiled
unmodified against oracle/shim). Run in the container that has /root/reference:

    python tests/golden/make_golden.py

Outputs
  tests/golden/tex_320x240_s11.pgm          input image (textured(320,240,11))
  tests/golden/tex_320x240_s11.ref.npz      every per-detection record of the reference (oracle.DET_DTYPE)
  tests/golden/tex_320x240_s11.hesaff.sift  the reference CLI's output file for that image
  tests/golden/summary.json                 counts + sha256 of the record bytes for larger images / other params
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tools.gen_textured import textured, write_pgm  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")

CASES = [  # name, w, h, seed, param overrides
    ("tex_320x240_s11", 320, 240, 11, {}),
    ("tex_640x480_s1", 640, 480, 1, {}),
    ("tex_333x251_s7", 333, 251, 7, {}),
    ("tex_640x480_s1_S10_oct3", 640, 480, 1, {"number_of_scales": 10, "max_octaves": 3}),
    ("tex_512x512_s5_thr5_oct6", 512, 512, 5, {"threshold": 5.0, "max_octaves": 6}),
]


def main():
    oracle.build(ref=True)
    ref = oracle.load("ref")
    os.makedirs(G, exist_ok=True)
    summary = {}
    for name, w, h, seed, over in CASES:
        img = textured(w, h, seed)
        d = ref.detect(img.astype(np.float32), ref.default_params(**over))
        summary[name] = {
            "w": w, "h": h, "seed": seed, "params": over,
            "image_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
            "detections": int(len(d)), "affine": int(d["affine_ok"].sum()), "described": int(d["described"].sum()),
            "records_sha256": hashlib.sha256(d.tobytes()).hexdigest(),
        }
        print(name, summary[name]["detections"], summary[name]["described"])
        if name == "tex_320x240_s11":
            pgm = os.path.join(G, name + ".pgm")
            write_pgm(pgm, img)
            np.savez_compressed(os.path.join(G, name + ".ref.npz"), dets=d)
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "hesaff_ref"), pgm])
            os.replace(pgm + ".hesaff.sift", os.path.join(G, name + ".hesaff.sift"))
    with open(os.path.join(G, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
