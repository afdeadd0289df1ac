"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the oracle on the same
seeded inputs and against the committed golden vectors of the reference build.  Run on the B200 box:
    python -m pytest tests -m gpu -x -q
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from parity import compare_keypoints
from tools.gen_textured import read_pgm, textured

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hesaff_b200
    return hesaff_b200


def run(hb, images, **over):
    a = np.asarray(images)
    if a.ndim == 2:
        a = a[None]
    par = hb.HessianAffineParams(**over)
    det = hb.AffineHessianDetector(par, device=0, max_width=a.shape[2], max_height=a.shape[1], max_batch=a.shape[0])
    det.detectPyramidKeypoints(a)
    return det


def oparams(orc, over):
    return orc.default_params(**over)


@pytest.mark.parametrize("w,h,seed,over", [
    (333, 251, 7, {}),
    (256, 192, 8, {"number_of_scales": 10, "max_octaves": 2}),
    (130, 67, 9, {}),
])
def test_pyramid_planes_bit_exact(hb, port_oracle, w, h, seed, over):
    """Blur planes L[i], response planes R[i] and the decimated next-octave seed, every octave and level."""
    img = textured(w, h, seed)
    det = run(hb, img, **over)
    p = oparams(port_oracle, over)
    first = port_oracle.first_level(img.astype(np.float32), p)
    n_oct, n_lvl, sizes = det.geometry()
    assert n_lvl == p.number_of_scales + 2
    for o in range(n_oct):
        assert sizes[o] == first.shape
        L, R, nxt = port_oracle.octave_planes(first, p)
        for l in range(n_lvl):
            gl = det.plane(0, o, l, "L")
            assert np.array_equal(gl, L[l]), (o, l, np.abs(gl - L[l]).max())
            gr = det.plane(0, o, l, "R")
            assert np.array_equal(gr, R[l]), (o, l, np.abs(gr - R[l]).max())
        first = nxt
    det.close()


CASES = [
    (320, 240, 11, {}),
    (333, 251, 7, {}),
    (640, 480, 1, {}),
    (640, 480, 1, {"number_of_scales": 10, "max_octaves": 3}),
    (512, 512, 5, {"threshold": 5.0, "max_octaves": 6}),
    (97, 131, 22, {}),
    (40, 30, 25, {}),
    (320, 240, 11, {"number_of_scales": 1}),     # second incremental blur: 35 taps -> the any-tap kernels (helpers.cpp:283-289)
    (320, 240, 11, {"desc_factor": 2.0}),        # imageToPatchScale <= 0.4: normalizeAffine samples 41x41 directly (affine.cpp:135-142)
]


@pytest.mark.parametrize("w,h,seed,over", CASES)
def test_detections_match_oracle(hb, port_oracle, w, h, seed, over):
    img = textured(w, h, seed)
    det = run(hb, img, **over)
    got = det.detections()
    want = port_oracle.detect(img.astype(np.float32), oparams(port_oracle, over))
    # detection stage: integer/branch logic on bit-exact planes -> identical set, order and values
    assert len(got) == len(want) == int(det.n_detected[0])
    for f in ("x", "y", "pd", "type", "response"):
        assert np.array_equal(got[f], want[f]), f
    assert np.allclose(got["s"], want["s"], rtol=3e-7, atol=0)          # 2^t: <= 1 ulp
    # affine stage: fp32 sums are reduced in a different order -> tiny drift, rare decision flips
    slack = max(1, int(0.005 * len(want)))      # decision flips allowed: 0.5 % (at least one)
    same = got["affine_ok"] == want["affine_ok"]
    assert (~same).sum() <= slack, (~same).sum()
    both = same & (want["affine_ok"] == 1)
    for f in ("u11", "u12", "u21", "u22", "a11", "a12", "a21", "a22"):
        assert np.allclose(got[f][both], want[f][both], rtol=0, atol=2e-4), f
    assert (got["iters"][both] != want["iters"][both]).sum() <= slack
    assert (got["described"] != want["described"]).sum() <= slack
    # final records, north_star tolerances
    kg = det.keys()
    kw = want[want["described"] == 1]
    st = compare_keypoints(kg, kw, mr_size=det.par.desc_factor)
    assert st["aligned"] >= len(kw) - slack and st["within_tol_frac"] * len(kw) >= len(kw) - slack, st
    assert int(det.n_described[0]) == len(kg)
    assert abs(len(kg) - len(kw)) <= slack, st
    det.close()


def test_golden_records_of_reference_build(hb):
    """Against tests/golden/ (written by the reference's own sources, tests/golden/make_golden.py)."""
    img = read_pgm(os.path.join(GOLDEN, "tex_320x240_s11.pgm"))
    want = np.load(os.path.join(GOLDEN, "tex_320x240_s11.ref.npz"))["dets"]
    det = run(hb, img)
    got = det.detections()
    assert len(got) == len(want)
    for f in ("x", "y", "pd", "type", "response"):
        assert np.array_equal(got[f], want[f]), f
    st = compare_keypoints(det.keys(), want[want["described"] == 1])
    assert st["within_tol_frac"] >= 0.995, st
    summ = json.load(open(os.path.join(GOLDEN, "summary.json")))["tex_320x240_s11"]
    assert len(got) == summ["detections"] and abs(len(det.keys()) - summ["described"]) <= 2
    det.close()


def test_patches_match_oracle(hb, port_oracle):
    """normalizeAffine output (41x41 patch before photometric normalisation) for the GPU's own (x,y,s,A)."""
    img = textured(400, 300, 33)
    det = run(hb, img)
    k = det.keys()
    P = det.patches(normalized=False)
    PN = det.patches(normalized=True)
    assert len(P) == len(k) > 500
    f = img.astype(np.float32)
    idx = np.linspace(0, len(k) - 1, 160).astype(int)
    big = np.argsort(-k["s"])[:24]           # exercise the MEDIUM/LARGE source-patch bins
    nbad = 0
    for i in np.unique(np.concatenate([idx, big])):
        A = [k["a11"][i], k["a12"][i], k["a21"][i], k["a22"][i]]
        rej, patch = port_oracle.normalize_affine(f, float(k["x"][i]), float(k["y"][i]), float(k["s"][i]), A)
        assert not rej
        assert np.array_equal(P[i], patch), (i, float(k["s"][i]), np.abs(P[i] - patch).max())
        desc, pn = port_oracle.sift(patch)
        assert np.allclose(PN[i], pn, rtol=0, atol=2e-3), (i, np.abs(PN[i] - pn).max())
        d = np.abs(desc.astype(int) - k["desc"][i].astype(int))
        nbad += int(d.max() > 1)
    assert nbad <= 2
    det.close()


def test_large_patches(hb, port_oracle):
    """A smooth image gives few, large-scale keypoints: source patches well beyond the shared-memory bins."""
    rng = np.random.default_rng(5)
    import cv2
    n = rng.standard_normal((600, 800)).astype(np.float32)
    g = cv2.GaussianBlur(n, (0, 0), 9.0)
    img = np.clip(128 + 60 * g / g.std(), 0, 255).astype(np.uint8)
    det = run(hb, img)
    got = det.detections()
    want = port_oracle.detect(img.astype(np.float32))
    assert len(got) == len(want) > 20
    kw = want[want["described"] == 1]
    assert (np.ceil(kw["s"] * det.par.desc_factor) * 2 + 3 > 95).sum() >= 5, "no large patches in this test image"
    st = compare_keypoints(det.keys(), kw)
    assert st["within_tol_frac"] >= 0.99, st
    k = det.keys()
    P = det.patches(normalized=False)
    f = img.astype(np.float32)
    for i in np.argsort(-k["s"])[:12]:
        A = [k["a11"][i], k["a12"][i], k["a21"][i], k["a22"][i]]
        rej, patch = port_oracle.normalize_affine(f, float(k["x"][i]), float(k["y"][i]), float(k["s"][i]), A)
        assert not rej and np.array_equal(P[i], patch), (i, float(k["s"][i]), np.abs(P[i] - patch).max())
    # the float-source kernels (fp32 / colour input) sample the float image instead of the u8 copy: same records
    detf = run(hb, f)
    assert detf.keys().tobytes() == k.tobytes()
    detf.close()
    det.close()


def test_batch_equals_single_image_runs_and_is_deterministic(hb):
    imgs = np.stack([textured(320, 240, s) for s in (41, 42, 43, 41)])
    det = run(hb, imgs)
    k = det.keys()
    o = det.offsets()
    assert np.array_equal(det.n_detected[0], det.n_detected[3]) and det.n_described[0] == det.n_described[3]
    assert k[o[0]:o[1]].tobytes() == k[o[3]:o[4]].tobytes()          # identical images -> identical records
    for i in range(3):
        d1 = run(hb, imgs[i])
        assert d1.keys().tobytes() == k[o[i]:o[i + 1]].tobytes()
        d1.close()
    det.detectPyramidKeypoints(imgs)                                    # run-to-run determinism
    assert det.keys().tobytes() == k.tobytes()
    det.close()


def test_chunked_batch_equals_one_chunk(hb):
    imgs = np.stack([textured(200, 150, 50 + s) for s in range(5)])
    a = run(hb, imgs)
    par = hb.HessianAffineParams()
    b = hb.AffineHessianDetector(par, 0, 200, 150, max_batch=2)          # 3 chunks: 2 + 2 + 1
    b.detectPyramidKeypoints(imgs)
    assert np.array_equal(a.n_detected, b.n_detected) and np.array_equal(a.n_described, b.n_described)
    assert a.keys().tobytes() == b.keys().tobytes()
    assert np.array_equal(a.ellipses(), b.ellipses())
    a.close(); b.close()


def test_f32_input_and_device_input(hb):
    import torch
    img = textured(320, 240, 61)
    a = run(hb, img)
    b = run(hb, img.astype(np.float32))
    assert a.keys().tobytes() == b.keys().tobytes()
    t = torch.from_numpy(img).cuda()
    c = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 320, 240, 1)
    c.detectPyramidKeypoints(t)
    assert a.keys().tobytes() == c.keys().tobytes()
    # pitched host input
    wide = np.zeros((240, 352), np.uint8)
    wide[:, :320] = img
    d = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 320, 240, 1)
    rc = hb.lib().hesaff_detect_u8(d._h, wide.ctypes.data, 1, 320, 240, 352, 352 * 240, 0, None)
    assert rc == 0
    d.n_images = 1
    assert a.keys().tobytes() == d.keys().tobytes()
    for x in (a, b, c, d):
        x.close()


def test_edge_cases(hb, port_oracle):
    # loop never runs (pyramid.cpp:284: rows > 12 && cols > 12)
    d = run(hb, textured(64, 12, 27))
    assert d.total() == 0 and int(d.n_detected[0]) == 0
    d.close()
    # exactly one octave
    img = textured(13, 13, 26)
    d = run(hb, img)
    assert int(d.n_detected[0]) == len(port_oracle.detect(img.astype(np.float32)))
    d.close()
    # constant image: no extrema, no NaNs escaping
    d = run(hb, np.full((100, 120), 77, np.uint8))
    assert d.total() == 0
    d.close()
    # empty batch
    det = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 64, 64, 2)
    det.detectPyramidKeypoints(np.zeros((0, 64, 64), np.uint8))
    assert det.total() == 0
    det.close()


def test_capacity_overflow_is_reported(hb):
    img = textured(320, 240, 71)
    det = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 320, 240, 1, max_candidates_per_image=64)
    with pytest.raises(hb.HesaffError, match="overflow"):
        det.detectPyramidKeypoints(img)
    det.close()


def test_ellipses_and_sift_file(hb, tmp_path):
    img = read_pgm(os.path.join(GOLDEN, "tex_320x240_s11.pgm"))
    det = run(hb, img)
    k = det.keys()
    from parity import ellipse
    e = det.ellipses()
    assert np.allclose(e, ellipse(k), rtol=2e-6, atol=1e-9)
    path = str(tmp_path / "out.hesaff.sift")
    assert det.exportKeypoints(path) == len(k)
    lines = open(path).read().split("\n")
    ref = open(os.path.join(GOLDEN, "tex_320x240_s11.hesaff.sift")).read().split("\n")
    assert lines[0] == ref[0] == "128"
    assert abs(int(lines[1]) - int(ref[1])) <= 2
    if int(lines[1]) == int(ref[1]):
        # same text except where a 6-digit value or a descriptor byte sits on a rounding edge
        want = np.array([[float(t) for t in ln.split()] for ln in ref[2:2 + len(k)]])
        have = np.array([[float(t) for t in ln.split()] for ln in lines[2:2 + len(k)]])
        # 6-digit text; the reference's float SVD vs the closed form: compare on the scale of the ellipse matrix
        scale = np.abs(want[:, 2:5]).max(1, keepdims=True)
        assert (np.abs(have[:, 2:5] - want[:, 2:5]) / scale).max() <= 1e-4
        assert np.allclose(have[:, :2], want[:, :2], rtol=1e-5, atol=0)
        assert np.abs(have[:, 5:] - want[:, 5:]).max() <= 2 and (have[:, 5:] == want[:, 5:]).mean() > 0.99
    rows = np.array([[float(t) for t in ln.split()] for ln in lines[2:2 + len(k)]])
    assert np.allclose(rows[:, :5], e, rtol=6e-6, atol=1e-9)
    det.close()


def test_host_cli_drop_in(hb, tmp_path):
    """The C++ host (`hesaff <image>`): same stdout line, same <image>.hesaff.sift format as hesaff.cpp:133-180."""
    import shutil
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "hesaff_b200", "host", "hesaff")
    if not os.path.exists(exe):
        pytest.skip("host CLI not built")
    img = str(tmp_path / "tex.pgm")
    shutil.copy(os.path.join(GOLDEN, "tex_320x240_s11.pgm"), img)
    out = subprocess.run([exe, img], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    summ = json.load(open(os.path.join(GOLDEN, "summary.json")))["tex_320x240_s11"]
    words = out.stdout.split()
    assert words[0] == "Detected" and int(words[1]) == summ["detections"] and words[2] == "keypoints"
    assert abs(int(words[4]) - summ["described"]) <= 2 and words[5:8] == ["affine", "shapes", "in"]
    got = open(img + ".hesaff.sift").read().split("\n")
    ref = open(os.path.join(GOLDEN, "tex_320x240_s11.hesaff.sift")).read().split("\n")
    assert got[0] == "128" and abs(int(got[1]) - int(ref[1])) <= 2
    if got[1] == ref[1]:
        a = np.array([[float(t) for t in ln.split()] for ln in got[2:2 + int(got[1])]])
        b = np.array([[float(t) for t in ln.split()] for ln in ref[2:2 + int(ref[1])]])
        assert np.allclose(a[:, :2], b[:, :2], rtol=1e-5, atol=0)
        assert np.abs(a[:, 5:] - b[:, 5:]).max() <= 2
    # colour input (P6) goes through the float path with gray = (R+G+B)/3.0f
    g = read_pgm(img)
    ppm = str(tmp_path / "tex.ppm")
    with open(ppm, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (g.shape[1], g.shape[0]))
        f.write(np.repeat(g[:, :, None], 3, 2).tobytes())
    out2 = subprocess.run([exe, ppm], capture_output=True, text=True, timeout=120)
    assert out2.returncode == 0 and out2.stdout.split()[1] == words[1]
    assert open(ppm + ".hesaff.sift").read() == open(img + ".hesaff.sift").read()
    # unreadable file: like the reference, an empty result and exit code 0
    out3 = subprocess.run([exe, str(tmp_path / "missing.pgm")], capture_output=True, text=True, timeout=60)
    assert out3.returncode == 0 and open(str(tmp_path / "missing.pgm") + ".hesaff.sift").read() == "128\n0\n"


@pytest.mark.parametrize("w,h,seed,over,crop", [
    (3840, 2160, 3, {"number_of_scales": 10, "max_octaves": 3}, (1408, 704, 1024, 768)),     # BASELINE configs[1]
    (4096, 4096, 5, {"threshold": 5.0, "max_octaves": 6}, (2048, 1024, 1024, 1024)),          # BASELINE configs[4]
])
def test_full_size_configs_match_oracle_on_an_interior_crop(hb, port_oracle, w, h, seed, over, crop):
    """Full BASELINE sizes through a size-independent property: every stage has finite support, so keypoints of the low
    octaves that lie well inside an octave-aligned crop are the same whether the path runs on the full frame (GPU) or
    on the crop alone (oracle, seconds).  Also: identical frames in one batch give identical records."""
    img = textured(w, h, seed)
    det = run(hb, np.stack([img, img]), **over)
    k = det.keys()
    o = det.offsets()
    assert k[o[0]:o[1]].tobytes() == k[o[1]:o[2]].tobytes()
    k = k[o[0]:o[1]]
    assert 10000 * (w * h / 1e6) < len(k) < 60000 * (w * h / 1e6)         # textured images: ~16-30 k keypoints / Mpix
    x0, y0, cw, ch = crop
    assert x0 % 32 == 0 and y0 % 32 == 0
    want = port_oracle.detect(img[y0:y0 + ch, x0:x0 + cw].astype(np.float32), oparams(port_oracle, over))
    want = want[(want["described"] == 1) & (want["pd"] <= 4)]
    m = 320                                                                   # > blur + SMM + patch support at pd <= 4
    inside = lambda x, y: (x > m) & (x < cw - m) & (y > m) & (y < ch - m)     # noqa: E731
    want = want[inside(want["x"], want["y"])]
    got = k[inside(k["x"] - x0, k["y"] - y0)].copy()
    got["x"] -= x0
    got["y"] -= y0
    # low octaves only (the GPU record has no pd; scale bounds it: s < sigma_max * 4 for pd <= 4)
    from scipy.spatial import cKDTree
    t = cKDTree(np.stack([got["x"], got["y"], got["s"]], 1))
    d, j = t.query(np.stack([want["x"], want["y"], want["s"]], 1))
    assert len(want) > 1000
    matched = d < 2e-3
    assert matched.mean() >= 0.995, matched.mean()
    st = compare_keypoints(got[j[matched]], want[matched], mr_size=det.par.desc_factor)
    assert st["within_tol_frac"] >= 0.995, st
    det.close()


def test_gpu_text_export_equals_host_writer(hb, tmp_path):
    """the .hesaff.sift text formatted on the GPU is byte-identical to the host ostream writer
    (exportKeypoints, hesaff.cpp:107-130) on the same keypoints, for every image of a batch; binary sidecar round trip."""
    imgs = np.stack([textured(320, 240, 11), textured(320, 240, 12), np.zeros((240, 320), np.uint8)])
    det = run(hb, imgs)
    k, off = det.keys(), det.offsets()
    for i in range(len(imgs)):
        host, gpu = str(tmp_path / ("h%d.sift" % i)), str(tmp_path / ("g%d.sift" % i))
        assert det.exportKeypoints(host, image=i, on_host=True) == off[i + 1] - off[i]
        assert det.exportKeypoints(gpu, image=i) == off[i + 1] - off[i]
        a, b = open(host, "rb").read(), open(gpu, "rb").read()
        assert a == b, "image %d: GPU text differs from the host writer" % i
        assert det.siftText(i) == a
    assert open(str(tmp_path / "g2.sift")).read() == "128\n0\n"
    bpath = str(tmp_path / "k.bin")
    det.exportKeypointsBinary(bpath, image=1)
    raw = open(bpath, "rb").read()
    assert raw[:8] == b"HESAFFB1" and len(raw) == 24 + 164 * (off[2] - off[1])
    back = np.frombuffer(raw[24:], hb.KEYPOINT_DTYPE)
    assert np.array_equal(back, k[off[1]:off[2]])
    n = hb.api.C.c_size_t()
    out = np.zeros(off[2] - off[1], hb.KEYPOINT_DTYPE)
    assert hb.lib().hesaff_read_keypoints_binary(bpath.encode(), out.ctypes.data, len(out), hb.api.C.byref(n)) == 0
    assert n.value == len(out) and np.array_equal(out, back)
    det.close()


def test_gpu_float_formatter_is_exact_percent_g(hb):
    """The device formatter against the host's correctly rounded "%g" (what ostream << float prints): random bit
    patterns over all finite magnitudes below 2^63, decimal ties, the %e/%f switch-over points, denormals, zeros."""
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 0x5F000000, 200000, dtype=np.uint32) | (rng.integers(0, 2, 200000, dtype=np.uint32) << 31)
    vals = [bits.view(np.float32)]
    vals.append(np.float32([0.0, -0.0, 1.0, 0.5, 123456.5, 123457.5, 999999.5, 999999.4, 1e6, 1234565.0, 1234575.0, 1e-4, 9.99999e-5,
                            9.999995e-5, 0.0001, 0.00012345675, 1e-5, 1.4e-45, 1.17549435e-38, 3.4e-39, 2.5, 0.125, 100000.0,
                            99999.95, 99999.94, 12345.675, 8388608.0, 16777216.0, 9.2e18, 7.0, 10.0, 1e10, 1.5e-10, 65504.0]))
    vals.append((rng.random(50000) * 4096).astype(np.float32))                      # coordinates
    vals.append((10.0 ** rng.uniform(-9, 0, 50000)).astype(np.float32))             # ellipse entries
    vals.append((rng.integers(0, 2 ** 24, 20000) / 2.0).astype(np.float32))         # x.5 values: ties at 6+ digits
    v = np.concatenate(vals)
    det = hb.AffineHessianDetector(hb.HessianAffineParams(), device=0, max_width=64, max_height=64, max_batch=1)
    got = det.formatFloats(v)
    want = ["%g" % float(x) for x in v]
    bad = [(float(x), g, w) for x, g, w in zip(v, got, want) if g != w]
    assert not bad, bad[:10]
    # values the device path does not cover are flagged, not misprinted
    assert det.formatFloats(np.float32([np.inf, np.nan, 1e19])) == ["?", "?", "?"]
    det.close()


def test_rgb8_ingest_matches_host_gray_conversion(hb):
    """interleaved 8-bit colour input, gray = (float(c0)+c1+c2)/3.0f (hesaff.cpp:138-148) on the GPU,
    equals the float path fed with the host-side conversion; a gray image replicated to 3 channels equals the u8 path."""
    rng = np.random.default_rng(3)
    g = textured(320, 240, 71)
    rgb = np.stack([g, np.roll(g, 3, 1), np.roll(g, 5, 0)], -1)
    rgb = np.clip(rgb.astype(np.int32) + rng.integers(-20, 21, rgb.shape), 0, 255).astype(np.uint8)
    gray = (rgb[..., 0].astype(np.float32) + rgb[..., 1].astype(np.float32) + rgb[..., 2].astype(np.float32)) / np.float32(3.0)
    a = run(hb, gray)
    b = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 320, 240, 2)
    b.detectPyramidKeypoints(np.stack([rgb, np.repeat(g[:, :, None], 3, 2)]))
    kb, ob = b.keys(), b.offsets()
    assert a.keys().tobytes() == kb[ob[0]:ob[1]].tobytes()
    c = run(hb, g)
    assert c.keys().tobytes() == kb[ob[1]:ob[2]].tobytes()
    for x in (a, b, c):
        x.close()


def test_device_records_and_nccl_keypoint_gather(hb):
    """the records stay on the GPU (zero-copy view of the library's buffer) and go through the
    variable-size all-gather over nccl (world size 1 here; two ranks are covered by the gloo test in test_host.py)."""
    import torch
    import torch.distributed as dist
    from hesaff_b200 import shard
    imgs = np.stack([textured(320, 240, 81), textured(320, 240, 82), textured(320, 240, 83)])
    det = run(hb, imgs)
    keys = det.keys()
    dev = torch.device("cuda:0")
    rec = shard.device_records(torch, det, dev)
    assert rec.data_ptr() == det.keys_device_ptr() and tuple(rec.shape) == (len(keys), 164)
    assert rec.cpu().numpy().tobytes() == keys.tobytes()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29700 + os.getpid() % 200))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        counts = shard.all_gather_counts(dist, torch, np.stack([det.n_detected, det.n_described], 1), dev)
        assert np.array_equal(counts[:, 1], det.n_described)
        got = shard.all_gather_keypoints(dist, torch, rec, counts, shard.partition(3, 1), dev)
        assert got.cpu().numpy().tobytes() == keys.tobytes()
    finally:
        dist.destroy_process_group()
    det.close()


KNOB_WORKER = r'''
import hashlib, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import hesaff_b200 as hb
from tools.gen_textured import textured
imgs = np.stack([textured(480, 360, 90 + s) for s in range(6)])
det = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 480, 360, max_batch=2)      # 3 chunks over two lanes
det.detectPyramidKeypoints(imgs)
print(hashlib.sha256(det.keys().tobytes()).hexdigest(), det.n_detected.tolist(), det.n_described.tolist())
'''


def test_schedule_knobs_do_not_change_results(hb, tmp_path):
    """The describe launch plan (which bins run where, with how many CTAs), a plan that leaves bins out, and the
    front-end / describe overlap across chunk lanes are scheduling only: records are byte-identical."""
    import subprocess
    import sys
    script = tmp_path / "k.py"
    script.write_text(KNOB_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({}, {"HESAFF_OVERLAP": "0"}, {"HESAFF_PLAN": "T2,S2,D1,E1,M1,L1"}, {"HESAFF_PLAN": "S3;L2"}, {"HESAFF_NO_STAGE": "1"},
                {"HESAFF_PLAN": "L1,M2,E3,D4,S7,T9;T1", "HESAFF_CHUNK": "1"}):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, str(script), root], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (env, r.stderr[-2000:])
        outs.append(r.stdout.strip().splitlines()[-1])
    assert len(set(outs)) == 1, outs


# ---- whole frames at the BASELINE sizes ---------------------------------------------------------------------
FULL_FRAMES = [
    ("full_1080p_s2", 1920, 1080, 2, {}),                                                   # BASELINE configs[2] frame
    ("full_4k_s3_S10_oct3", 3840, 2160, 3, {"number_of_scales": 10, "max_octaves": 3}),    # BASELINE configs[1]
    ("full_4096_s5_thr5_oct6", 4096, 4096, 5, {"threshold": 5.0, "max_octaves": 6}),        # BASELINE configs[4]
]


def _detection_hash(d):
    import hashlib
    h = hashlib.sha256()
    for f in ("x", "y", "pd", "type", "response"):
        h.update(np.ascontiguousarray(d[f]).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name,w,h,seed,over", FULL_FRAMES)
def test_full_frames_match_golden_of_the_reference_build(hb, name, w, h, seed, over):
    """The WHOLE frame against tests/golden/full_*.npz, written by the reference's own sources (make_golden.py): detection
    count, sha256 of every detection's bit-exact fields, and every STRIDE-th full record within north_star's tolerances."""
    summ = json.load(open(os.path.join(GOLDEN, "summary.json")))[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    img = textured(w, h, seed)
    import hashlib
    assert hashlib.sha256(img.tobytes()).hexdigest() == summ["image_sha256"]
    det = run(hb, img, **over)
    got = det.detections()
    assert len(got) == summ["detections"] == int(det.n_detected[0])
    assert _detection_hash(got) == summ["detection_fields_sha256"]
    slack = max(2, int(0.0005 * summ["detections"]))
    assert abs(int(got["affine_ok"].sum()) - summ["affine"]) <= slack
    assert abs(int(det.n_described[0]) - summ["described"]) <= slack
    idx, want = gold["index"], gold["dets"]
    sub = got[idx]
    for f in ("x", "y", "pd", "type", "response"):
        assert np.array_equal(sub[f], want[f]), f
    same = (sub["affine_ok"] == want["affine_ok"]) & (sub["described"] == want["described"])
    assert (~same).sum() <= max(1, int(0.002 * len(want)))
    both = same & (want["described"] == 1)
    st = compare_keypoints(sub[both], want[both], mr_size=det.par.desc_factor)
    assert st["aligned"] == int(both.sum()) and st["within_tol_frac"] >= 0.998, st
    assert st["desc_within_1_frac"] >= 0.995, st
    det.close()


def test_full_1080p_frame_against_the_reference_build_directly(hb, ref_oracle):
    """One hop instead of two: the CUDA path against oracle/_ref (the reference's own translation units, which travel to
    the GPU box as a built library) on a whole 1920x1080 frame, every record."""
    img = textured(1920, 1080, 2)
    want = ref_oracle.detect(img.astype(np.float32))
    det = run(hb, img)
    got = det.detections()
    assert len(got) == len(want) > 38000
    for f in ("x", "y", "pd", "type", "response"):
        assert np.array_equal(got[f], want[f]), f
    slack = int(0.0005 * len(want))
    assert (got["affine_ok"] != want["affine_ok"]).sum() <= slack
    assert (got["described"] != want["described"]).sum() <= slack
    kw = want[want["described"] == 1]
    assert len(kw) > 36000
    st = compare_keypoints(det.keys(), kw, mr_size=det.par.desc_factor)
    assert st["aligned"] >= len(kw) - slack and st["within_tol_frac"] * len(kw) >= len(kw) - 2 * slack, st
    det.close()


# ---- the steps either side of the path ----------------------------------------------------------------------
def test_pnm_files_go_to_the_gpu_as_they_are(hb):
    """hesaff_detect_pnm (8(f) rank 2): raw P5 / P6 file bytes in, header parsed on the host, pixels converted on the GPU;
    same records as the array entry points."""
    g = np.stack([textured(200, 150, 61), textured(200, 150, 62)])
    det = run(hb, g)
    want = det.keys().tobytes()
    p5 = [b"P5\n# a comment\n200 150\n255\n" + im.tobytes() for im in g]
    det.detectFiles(p5)
    assert det.keys().tobytes() == want and det.n_described.tolist() == [int(v) for v in det.n_described]
    p6 = [b"P6 200 150 255\n" + np.repeat(im[:, :, None], 3, 2).tobytes() for im in g]      # gray stored as colour
    det.detectFiles(p6)
    assert det.keys().tobytes() == want
    with pytest.raises(hb.HesaffError):
        det.detectFiles([p5[0], b"P5\n100 150\n255\n" + bytes(100 * 150)])                   # mixed sizes
    with pytest.raises(hb.HesaffError):
        det.detectFiles([p5[0][:-10]])                                                       # truncated payload
    det.close()


def test_descriptor_matcher_is_exact(hb):
    """hesaff_match_descriptors (8(f) rank 3, the consumer): nearest / second nearest by squared L2 over the 128 bytes
    equal a numpy brute force, ties to the lower index."""
    import torch
    from hesaff_b200 import shard
    a, b = textured(320, 240, 71), textured(320, 240, 71)
    b = np.roll(b, 3, axis=1)                                 # the same texture shifted: many true matches
    det = run(hb, np.stack([a, b]))
    keys = det.keys()
    off = det.offsets()
    dev = torch.device("cuda:0")
    rec = shard.device_records(torch, det, dev)
    q, d = rec[off[0]:off[1]], rec[off[1]:off[2]]
    idx, d1, d2 = (t.cpu().numpy() for t in shard.match_descriptors(torch, q, d, dev))
    qa = keys["desc"][off[0]:off[1]].astype(np.int64)
    da = keys["desc"][off[1]:off[2]].astype(np.int64)
    dist = ((qa[:, None, :] - da[None, :, :]) ** 2).sum(2)
    order = np.argsort(dist, axis=1, kind="stable")
    assert np.array_equal(idx, order[:, 0])
    assert np.array_equal(d1, dist[np.arange(len(qa)), order[:, 0]])
    assert np.array_equal(d2, dist[np.arange(len(qa)), order[:, 1]])
    # the shift is 3 px: most keypoints find their own copy (distance ratio test of Lowe)
    good = d1 < 0.36 * d2
    dx = keys["x"][off[1]:off[2]][idx[good]] - keys["x"][off[0]:off[1]][good]
    assert good.sum() > 300 and np.mean(np.abs(dx - 3) < 0.5) > 0.9
    i0, e1, _ = shard.match_descriptors(torch, q, rec[:0], dev)
    assert (i0.cpu().numpy() == -1).all() and (e1.cpu().numpy() == -1).all()
    det.close()


TWO_RANK_WORKER = r'''
import hashlib, json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch
import torch.distributed as dist
import hesaff_b200 as hb
from hesaff_b200 import shard
from tools.gen_textured import textured
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = 5                                                           # uneven blocks: 3 + 2
imgs = np.stack([textured(320, 240, 40 + s) for s in range(n)])
starts = shard.partition(n, world)
mine = imgs[starts[rank]:starts[rank + 1]]
det = hb.AffineHessianDetector(hb.HessianAffineParams(), rank, 320, 240, max_batch=len(mine))
det.detectPyramidKeypoints(mine)
counts = shard.all_gather_counts(dist, torch, np.stack([det.n_detected, det.n_described], 1), dev)
rec = shard.device_records(torch, det, dev)
allrec = shard.all_gather_keypoints(dist, torch, rec, counts, starts, dev)      # 164-byte records over NVLink
idx, d1, d2 = shard.match_descriptors(torch, rec, allrec, dev)                  # consumer: every local record finds itself
off = shard.global_offsets(counts)
self_index = np.arange(off[starts[rank]], off[starts[rank + 1]])
out = {"rank": rank, "counts": counts.tolist(), "sha": hashlib.sha256(allrec.cpu().numpy().tobytes()).hexdigest(),
       "self_match": bool((d1.cpu().numpy() == 0).all() and (idx.cpu().numpy() <= self_index).all()), "n_local": int(rec.shape[0])}
print("RESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_equal_one_gpu_and_gather_over_nccl(hb, tmp_path):
    """SURVEY App. C / 8(e),(f)3 on hardware: two ranks, each with its block of the batch on its own GPU; the concatenated
    records (variable-size all-gather over nccl) are byte-identical to the single-GPU run, on every rank."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(TWO_RANK_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29800 + os.getpid() % 100), str(script), root], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = [json.loads(ln.split("RESULT ", 1)[1]) for ln in r.stdout.splitlines() if "RESULT " in ln]
    assert sorted(x["rank"] for x in res) == [0, 1]
    imgs = np.stack([textured(320, 240, 40 + s) for s in range(5)])
    det = run(hb, imgs)
    import hashlib
    want = hashlib.sha256(det.keys().tobytes()).hexdigest()
    for x in res:
        assert x["sha"] == want and x["self_match"]
        assert [c[1] for c in x["counts"]] == det.n_described.tolist() and [c[0] for c in x["counts"]] == det.n_detected.tolist()
    assert sum(x["n_local"] for x in res) == det.total()
    det.close()


SANITIZER_WORKER = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import hesaff_b200 as hb
from tools.gen_textured import textured
img = textured(640, 480, 1)                       # BASELINE configs[0]
det = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 640, 480, max_batch=1)
det.detectPyramidKeypoints(img)
print("RESULT", int(det.n_detected[0]), int(det.n_described[0]), flush=True)
det.close()
'''


SANITIZER_WORKER_LARGE = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import cv2
import hesaff_b200 as hb
rng = np.random.default_rng(5)
g = cv2.GaussianBlur(rng.standard_normal((600, 800)).astype(np.float32), (0, 0), 9.0)
img = np.clip(128 + 60 * g / g.std(), 0, 255).astype(np.uint8)      # smooth: hundreds of source patches beyond 95 px
det = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 800, 600, max_batch=2)
for batch in (np.stack([img, img[::-1].copy()]), np.stack([img, img[::-1].copy()]).astype(np.float32)):   # u8- and float-source kernels
    det.detectPyramidKeypoints(batch)
    k = det.keys()
    P = 2 * np.ceil(k["s"] * det.par.desc_factor).astype(int) + 3
    print("LARGE", int((P > 95).sum()), flush=True)
det.close()
'''


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_on_large_patches(hb, tmp_path, tool):
    """The LARGE-bin kernel (shared-memory source boxes, lane-exchanged scratch stores) under compute-sanitizer; opt-in."""
    import shutil
    import subprocess
    import sys
    if os.environ.get("HESAFF_SANITIZER") != "1":
        pytest.skip("set HESAFF_SANITIZER=1 to run compute-sanitizer (slow)")
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(SANITIZER_WORKER_LARGE)
    r = subprocess.run([exe, "--tool", tool, "--error-exitcode", "77", sys.executable, str(script), root],
                       capture_output=True, text=True, timeout=3000)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    n = [int(ln.split()[1]) for ln in r.stdout.splitlines() if ln.startswith("LARGE")]
    assert len(n) == 2 and n[0] == n[1] > 300
    assert ("ERROR SUMMARY: 0 errors" in out) if tool == "memcheck" else ("RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in out), out[-2000:]


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_on_config0(hb, tmp_path, tool):
    """Hygiene, opt-in: HESAFF_SANITIZER=1 runs BASELINE configs[0] under compute-sanitizer.  memcheck:
    no out-of-bounds / misaligned access in any kernel; racecheck: no shared-memory hazard (the describe kernels alias
    their buffers between phases).  Minutes per tool, hence not in the default run."""
    import shutil
    import subprocess
    import sys
    if os.environ.get("HESAFF_SANITIZER") != "1":
        pytest.skip("set HESAFF_SANITIZER=1 to run compute-sanitizer (slow)")
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(SANITIZER_WORKER)
    r = subprocess.run([exe, "--tool", tool, "--error-exitcode", "77", sys.executable, str(script), root],
                       capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "RESULT 5596 4976" in r.stdout                      # SURVEY 8(c) checksum of this fixture
    out = r.stdout + r.stderr
    assert ("ERROR SUMMARY: 0 errors" in out) if tool == "memcheck" else ("RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in out), out[-2000:]


def test_host_cli_on_two_gpus(hb, tmp_path):
    """The C++ host with --gpus 2: contiguous blocks of files per GPU, one NCCL all-gather of the per-image counts; every
    output file equals the single-GPU tool's."""
    import subprocess
    import torch
    from tools.gen_textured import write_pgm
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "hesaff_b200", "host", "hesaff")
    if not os.path.exists(exe):
        pytest.skip("host CLI not built")
    files = []
    for s in range(3):
        f = str(tmp_path / ("img%d.pgm" % s))
        write_pgm(f, textured(320 + 16 * s, 240, 50 + s))      # different sizes: one context per file
        files.append(f)
    one = subprocess.run([exe] + files, capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stderr
    want = [open(f + ".hesaff.sift").read() for f in files]
    for f in files:
        os.remove(f + ".hesaff.sift")
    two = subprocess.run([exe, "--gpus", "2"] + files, capture_output=True, text=True, timeout=300)
    assert two.returncode == 0, two.stderr
    assert [open(f + ".hesaff.sift").read() for f in files] == want
    l1 = one.stdout.strip().splitlines()
    l2 = [ln for ln in two.stdout.strip().splitlines() if not ln.startswith("NCCL version")]     # NCCL_DEBUG=VERSION banner
    assert len(l2) == len(l1) + 1 and "counts all-gathered with NCCL" in l2[-1]
    assert [ln.split(" in ")[0] for ln in l2[:-1]] == [ln.split(" in ")[0] for ln in l1]
    tot = sum(int(ln.split()[1]) for ln in l1)
    assert int(l2[-1].split(": ")[1].split()[0]) == tot
