#!/usr/bin/env python
"""Synthetic "textured" test images: a sum of unit-variance band-limited
Gaussian noise fields at sigma = 1,2,4,8,16,32 px, rescaled to mean 128 / std 48 and clipped to u8.

    python tools/gen_textured.py W H SEED out.pgm

`textured(w, h, seed)` is also imported by tests and bench.py.  Uses cv2.GaussianBlur when cv2 is
importable (the generator SURVEY.md's probe fixtures were made with), else scipy.ndimage.
"""
import sys
import numpy as np


def _blur(n, sigma):
    try:
        import cv2
        cv2.setNumThreads(1)
        return cv2.GaussianBlur(n, (0, 0), sigma, borderType=cv2.BORDER_REFLECT_101)
    except ImportError:  # pragma: no cover
        from scipy.ndimage import gaussian_filter
        return gaussian_filter(n, sigma, mode="mirror").astype(np.float32)


def textured(w, h, seed):
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.float64)
    for sigma in (1, 2, 4, 8, 16, 32):
        n = rng.standard_normal((h, w)).astype(np.float32)
        g = _blur(n, sigma)
        acc += g / g.std()
    z = (acc - acc.mean()) / acc.std()
    return np.clip(128 + 48 * z, 0, 255).astype(np.uint8)


def write_pgm(path, img):
    h, w = img.shape
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end])
        pos = end
    pos += 1
    assert toks[0] == b"P5" and toks[3] == b"255"
    w, h = int(toks[1]), int(toks[2])
    return np.frombuffer(data, np.uint8, w * h, pos).reshape(h, w).copy()


if __name__ == "__main__":
    w, h, seed, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    write_pgm(out, textured(w, h, seed))
