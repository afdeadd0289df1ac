#!/usr/bin/env python
"""Generate tests/golden/ from the REFERENCE build (oracle/_ref: perdoch/hesaff sources compiled
unmodified against oracle/shim). Run in the container that has /root/reference:

    python tests/golden/make_golden.py

Outputs
  tests/golden/tex_320x240_s11.pgm          input image (textured(320,240,11))
  tests/golden/tex_320x240_s11.ref.npz      every per-detection record of the reference (oracle.DET_DTYPE)
  tests/golden/tex_320x240_s11.hesaff.sift  the reference CLI's output file for that image
  tests/golden/summary.json                 counts + sha256 of the record bytes for larger images / other params
  tests/golden/full_<name>.npz              BASELINE-size frames (1080p default; 4K S=10 3 octaves; 4096^2 thr 5 6 octaves):
                                            every STRIDE-th per-detection record of the reference, whole-frame counts and
                                            the sha256 of the bit-exact detection fields (x, y, pd, type, response) of ALL
                                            detections -- what tests/test_gpu_parity.py::test_full_frames_* compare with
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
    # desc_factor 2: most keypoints have imageToPatchScale <= 0.4, the direct-sampling branch of normalizeAffine (affine.cpp:135-142)
    ("tex_320x240_s11_df2", 320, 240, 11, {"desc_factor": 2.0}),
    ("tex_320x240_s11_S1", 320, 240, 11, {"number_of_scales": 1}),
]


FULL = [  # name, w, h, seed, overrides, record stride      (SURVEY.md 8(c) fixtures / BASELINE.json configs 2, 1, 4)
    ("full_1080p_s2", 1920, 1080, 2, {}, 16),
    ("full_4k_s3_S10_oct3", 3840, 2160, 3, {"number_of_scales": 10, "max_octaves": 3}, 64),
    ("full_4096_s5_thr5_oct6", 4096, 4096, 5, {"threshold": 5.0, "max_octaves": 6}, 64),
]


def detection_hash(d):
    """sha256 over the fields the CUDA path reproduces bit for bit, in reference order"""
    h = hashlib.sha256()
    for f in ("x", "y", "pd", "type", "response"):
        h.update(np.ascontiguousarray(d[f]).tobytes())
    return h.hexdigest()


def full_frames(ref, summary):
    for name, w, h, seed, over, stride in FULL:
        img = textured(w, h, seed)
        d = ref.detect(img.astype(np.float32), ref.default_params(**over))
        summary[name] = {
            "w": w, "h": h, "seed": seed, "params": over, "stride": stride,
            "image_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
            "detections": int(len(d)), "affine": int(d["affine_ok"].sum()), "described": int(d["described"].sum()),
            "detection_fields_sha256": detection_hash(d),
            "records_sha256": hashlib.sha256(d.tobytes()).hexdigest(),
        }
        print(name, summary[name]["detections"], summary[name]["described"], flush=True)
        np.savez_compressed(os.path.join(G, name + ".npz"), index=np.arange(0, len(d), stride), dets=d[::stride])


def main():
    oracle.build(ref=True)
    ref = oracle.load("ref")
    os.makedirs(G, exist_ok=True)
    summary = {}
    if "--no-full" in sys.argv and os.path.exists(os.path.join(G, "summary.json")):   # keep the full-frame entries
        summary = {k: v for k, v in json.load(open(os.path.join(G, "summary.json"))).items() if k.startswith("full_")}
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
    if "--no-full" not in sys.argv:
        full_frames(ref, summary)
    with open(os.path.join(G, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
