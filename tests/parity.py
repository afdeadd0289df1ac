"""Index alignment + tolerances used by the GPU parity tests, smoke() and bench.py's self-check.

north_star tolerances: (u, v, a, b, c) within 1e-3 and descriptor L2 <= 1e-2 (descriptor on the unit scale,
i.e. u8/512, siftdesc.cpp:110) per keypoint after index alignment.
"""
import numpy as np

MR_SIZE = np.float32(3.0) * np.sqrt(np.float32(3.0))


def ellipse(k, mr_size=MR_SIZE):
    """(u,v,a,b,c) of exportKeypoints (hesaff.cpp:115-125): E = (A A^T)^-1 / (mrSize*s)^2, float64."""
    a11, a12, a21, a22 = (k[n].astype(np.float64) for n in ("a11", "a12", "a21", "a22"))
    sc = np.float64(mr_size) * k["s"].astype(np.float64)
    p, q, r = a11 * a11 + a12 * a12, a11 * a21 + a12 * a22, a21 * a21 + a22 * a22
    det = p * r - q * q
    with np.errstate(all="ignore"):
        return np.stack([k["x"].astype(np.float64), k["y"].astype(np.float64), r / det / sc ** 2, -q / det / sc ** 2,
                         p / det / sc ** 2], 1)


def align(got, want):
    """Index pairs (i_got, i_want). Both are in reference order, so equal-length inputs with matching positions
    align one to one; otherwise match on the exact detection tuple (x, y, s) produced by localisation."""
    if len(got) == len(want) and np.array_equal(got["x"], want["x"]) and np.array_equal(got["y"], want["y"]):
        idx = np.arange(len(got))
        return idx, idx
    key = lambda a: {(float(x), float(y), float(pd if pd is not None else 0)): i  # noqa: E731
                     for i, (x, y, pd) in enumerate(zip(a["x"], a["y"], a["pd"] if "pd" in a.dtype.names else [None] * len(a)))}
    kg, kw = key(got), key(want)
    common = [k for k in kw if k in kg]
    if len(common) < 0.9 * len(kw):   # fall back to nearest neighbour on (x, y, s)
        from scipy.spatial import cKDTree
        t = cKDTree(np.stack([got["x"], got["y"], got["s"]], 1))
        d, j = t.query(np.stack([want["x"], want["y"], want["s"]], 1))
        ok = d < 1e-2
        return j[ok], np.nonzero(ok)[0]
    return np.array([kg[k] for k in common], np.int64), np.array([kw[k] for k in common], np.int64)


def compare_keypoints(got, want, mr_size=MR_SIZE):
    """got/want: structured arrays with x,y,s,a11..a22,desc (described keypoints only). Returns a stats dict."""
    ig, iw = align(got, want)
    g, w = got[ig], want[iw]
    eg, ew = ellipse(g, mr_size), ellipse(w, mr_size)
    duv = np.abs(eg[:, :2] - ew[:, :2]).max(1) if len(g) else np.zeros(0)
    dabc = np.abs(eg[:, 2:] - ew[:, 2:]).max(1) if len(g) else np.zeros(0)
    dl2 = np.sqrt(((g["desc"].astype(np.float64) - w["desc"].astype(np.float64)) ** 2).sum(1)) / 512.0 if len(g) else np.zeros(0)
    ok = (duv <= 1e-3) & (dabc <= 1e-3) & (dl2 <= 1e-2)
    return {
        "n_got": int(len(got)), "n_want": int(len(want)), "aligned": int(len(g)),
        "aligned_frac": float(len(g) / max(1, len(want))),
        "within_tol_frac": float(ok.sum() / max(1, len(want))),
        "max_duv": float(duv.max()) if len(g) else 0.0,
        "max_dabc": float(dabc.max()) if len(g) else 0.0,
        "max_desc_l2": float(dl2.max()) if len(g) else 0.0,
        "desc_identical_frac": float((dl2 == 0).mean()) if len(g) else 1.0,
        "desc_within_1_frac": float((np.abs(g["desc"].astype(int) - w["desc"].astype(int)).max(1) <= 1).mean()) if len(g) else 1.0,
    }
