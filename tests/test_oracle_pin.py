"""Pins the plain-C oracle (oracle/hesaff_oracle.c) to the reference:
  * bit for bit against oracle/_ref (the reference's own sources, compiled unmodified), per stage and end to end;
  * against the golden vectors under tests/golden/ that tests/golden/make_golden.py wrote from oracle/_ref.
The reference ships no tests or fixtures of its own, so these are the pins."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from tools.gen_textured import read_pgm, textured

SUMMARY = json.load(open(os.path.join(GOLDEN, "summary.json")))


def _img(case):
    c = SUMMARY[case]
    img = textured(c["w"], c["h"], c["seed"])
    if hashlib.sha256(img.tobytes()).hexdigest() != c["image_sha256"]:
        pytest.skip("generator produced different pixels here (different cv2 build?)")
    return img.astype(np.float32), c


@pytest.mark.parametrize("case", sorted(SUMMARY))
def test_port_matches_golden_summary(case, port_oracle):
    img, c = _img(case)
    d = port_oracle.detect(img, port_oracle.default_params(**c["params"]))
    assert len(d) == c["detections"]
    assert int(d["affine_ok"].sum()) == c["affine"]
    assert int(d["described"].sum()) == c["described"]
    assert hashlib.sha256(d.tobytes()).hexdigest() == c["records_sha256"]


def test_port_matches_golden_records(port_oracle):
    img = read_pgm(os.path.join(GOLDEN, "tex_320x240_s11.pgm")).astype(np.float32)
    want = np.load(os.path.join(GOLDEN, "tex_320x240_s11.ref.npz"))["dets"]
    got = port_oracle.detect(img)
    assert got.dtype == want.dtype
    assert got.tobytes() == want.tobytes()


def test_golden_sift_file_consistent_with_records():
    """The reference CLI's text output and the binary records describe the same keypoints."""
    want = np.load(os.path.join(GOLDEN, "tex_320x240_s11.ref.npz"))["dets"]
    k = want[want["described"] == 1]
    lines = open(os.path.join(GOLDEN, "tex_320x240_s11.hesaff.sift")).read().split("\n")
    assert lines[0] == "128" and int(lines[1]) == len(k)
    rows = np.array([[float(t) for t in ln.split()] for ln in lines[2:2 + len(k)]])
    assert rows.shape == (len(k), 133)
    assert np.allclose(rows[:, 0], k["x"], rtol=6e-6, atol=0) and np.allclose(rows[:, 1], k["y"], rtol=6e-6, atol=0)
    assert np.array_equal(rows[:, 5:].astype(np.uint8), k["desc"])
    # ellipse: E = (A A^T)^-1 / (mrSize*s)^2   (hesaff.cpp:115-125)
    sc = np.float64(3.0 * np.sqrt(3.0)) * k["s"].astype(np.float64)
    a11, a21, a22 = (k[n].astype(np.float64) for n in ("a11", "a21", "a22"))
    # A = [a11 0; a21 a22] -> A A^T = [a11^2, a11 a21; a11 a21, a21^2+a22^2]
    p, q, r = a11 * a11, a11 * a21, a21 * a21 + a22 * a22
    det = p * r - q * q
    E = np.stack([r / det, -q / det, p / det], 1) / (sc * sc)[:, None]
    assert np.allclose(rows[:, 2:5], E, rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("w,h,seed,over", [
    (200, 150, 21, {}),
    (97, 131, 22, {}),
    (160, 120, 23, {"number_of_scales": 5}),
    (160, 120, 24, {"threshold": 3.0, "max_octaves": 2}),
    (40, 30, 25, {}),
    (13, 13, 26, {}),   # rows>12 && cols>12 just true: one octave
    (12, 64, 27, {}),   # loop never runs (pyramid.cpp:284)
])
def test_port_equals_reference_build_end_to_end(w, h, seed, over, port_oracle, ref_oracle):
    img = textured(w, h, seed).astype(np.float32)
    a = ref_oracle.detect(img, ref_oracle.default_params(**over))
    b = port_oracle.detect(img, port_oracle.default_params(**over))
    assert a.tobytes() == b.tobytes()


def test_port_equals_reference_build_per_stage(port_oracle, ref_oracle):
    img = textured(192, 144, 31).astype(np.float32)
    f_a, f_b = ref_oracle.first_level(img), port_oracle.first_level(img)
    assert np.array_equal(f_a, f_b)
    La, Ra, na = ref_oracle.octave_planes(f_a)
    Lb, Rb, nb = port_oracle.octave_planes(f_b)
    assert np.array_equal(La, Lb) and np.array_equal(Ra, Rb) and np.array_equal(na, nb)
    assert np.array_equal(ref_oracle.hessian_response(img, 2.56), port_oracle.hessian_response(img, 2.56))
    dets = ref_oracle.detect(img)
    assert len(dets) > 100
    n_aff = n_patch = 0
    for d in dets[:: max(1, len(dets) // 60)]:
        if d["pd"] != 1.0:
            continue
        # the affine iteration runs on some blur level; any plane exercises the code identically
        for plane in (La[0], La[2]):
            oa = ref_oracle.find_affine_shape(plane, d["x"], d["y"], d["s"], 1.0)
            ob = port_oracle.find_affine_shape(plane, d["x"], d["y"], d["s"], 1.0)
            assert oa[0] == ob[0]
            if oa[0]:
                n_aff += 1
                assert np.array_equal(oa[1], ob[1]) and oa[2] == ob[2]
                A = ref_oracle.rectify(oa[1])
                assert np.array_equal(A, port_oracle.rectify(ob[1]))
                ra, pa = ref_oracle.normalize_affine(img, d["x"], d["y"], d["s"], A)
                rb, pb = port_oracle.normalize_affine(img, d["x"], d["y"], d["s"], A)
                assert ra == rb
                if not ra:
                    n_patch += 1
                    assert np.array_equal(pa, pb)
                    da, qa = ref_oracle.sift(pa)
                    db, qb = port_oracle.sift(pb)
                    assert np.array_equal(da, db) and np.array_equal(qa, qb)
    assert n_aff > 20 and n_patch > 10


def test_default_params_equal_reference_structs(port_oracle, ref_oracle):
    a, b = ref_oracle.default_params(), port_oracle.default_params()
    for name, _ in a._fields_:
        assert getattr(a, name) == getattr(b, name), name
    assert abs(a.threshold - 16.0 / 3.0) < 1e-6 and a.max_iter == 16 and a.patch_size == 41
    assert a.number_of_scales == 3 and a.border == 5 and a.smm_window_size == 19
