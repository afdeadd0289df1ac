"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for the two implementations of oracle_api.h:
    load("port") -> oracle/libhesaff_oracle.so   (plain-C restatement, hesaff_oracle.c)
    load("ref")  -> oracle/_ref/libhesaff_ref.so (the reference's own sources + oracle/shim)
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Params(C.Structure):
    _fields_ = [("threshold", C.c_float), ("max_iter", C.c_int), ("desc_factor", C.c_float), ("patch_size", C.c_int),
                ("number_of_scales", C.c_int), ("initial_sigma", C.c_float), ("edge_eigenvalue_ratio", C.c_float),
                ("border", C.c_int), ("convergence_threshold", C.c_float), ("smm_window_size", C.c_int),
                ("max_octaves", C.c_int)]


DET_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("s", "f4"), ("pd", "f4"), ("type", "i4"), ("response", "f4"),
                      ("affine_ok", "i4"), ("u11", "f4"), ("u12", "f4"), ("u21", "f4"), ("u22", "f4"), ("iters", "i4"),
                      ("described", "i4"), ("a11", "f4"), ("a12", "f4"), ("a21", "f4"), ("a22", "f4"),
                      ("desc", "u1", (128,))])
assert DET_DTYPE.itemsize == 196

_FP = C.POINTER(C.c_float)


def _fp(a):
    return a.ctypes.data_as(_FP)


class Oracle:
    def __init__(self, path):
        self.path = path
        self.lib = L = C.CDLL(path)
        L.orc_name.restype = C.c_char_p
        L.orc_detect.restype = C.c_int
        L.orc_detect.argtypes = [_FP, C.c_int, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_gaussian_blur.argtypes = [_FP, C.c_int, C.c_int, C.c_float, _FP]
        L.orc_hessian_response.argtypes = [_FP, C.c_int, C.c_int, C.c_float, _FP]
        L.orc_octave_planes.argtypes = [_FP, C.c_int, C.c_int, C.POINTER(Params), _FP, _FP, _FP]
        L.orc_first_level.argtypes = [_FP, C.c_int, C.c_int, C.POINTER(Params), _FP]
        L.orc_find_affine_shape.restype = C.c_int
        L.orc_find_affine_shape.argtypes = [_FP, C.c_int, C.c_int, C.POINTER(Params)] + [C.c_float] * 4 + [_FP, C.POINTER(C.c_int)]
        L.orc_rectify.argtypes = [_FP]
        L.orc_normalize_affine.restype = C.c_int
        L.orc_normalize_affine.argtypes = [_FP, C.c_int, C.c_int, C.POINTER(Params)] + [C.c_float] * 7 + [_FP]
        L.orc_sift.argtypes = [_FP, C.POINTER(Params), C.POINTER(C.c_ubyte)]
        self.name = L.orc_name().decode()

    def default_params(self, **kw):
        p = Params()
        self.lib.orc_default_params(C.byref(p))
        for k, v in kw.items():
            assert hasattr(p, k), k
            setattr(p, k, v)
        return p

    @staticmethod
    def _img(image):
        return np.ascontiguousarray(image, dtype=np.float32)

    def detect(self, image, params=None):
        """Full path; returns a structured array (DET_DTYPE), one row per detection, reference order."""
        img = self._img(image)
        p = params or self.default_params()
        out = C.c_void_p()
        n = self.lib.orc_detect(_fp(img), img.shape[0], img.shape[1], C.byref(p), C.byref(out))
        buf = C.string_at(out.value, n * DET_DTYPE.itemsize) if n else b""
        self.lib.orc_free(out)
        return np.frombuffer(buf, DET_DTYPE).copy()

    def gaussian_blur(self, image, sigma):
        img = self._img(image)
        dst = np.empty_like(img)
        self.lib.orc_gaussian_blur(_fp(img), img.shape[0], img.shape[1], sigma, _fp(dst))
        return dst

    def hessian_response(self, image, norm):
        img = self._img(image)
        dst = np.empty_like(img)
        self.lib.orc_hessian_response(_fp(img), img.shape[0], img.shape[1], norm, _fp(dst))
        return dst

    def first_level(self, image, params=None):
        img = self._img(image)
        p = params or self.default_params()
        dst = np.empty_like(img)
        self.lib.orc_first_level(_fp(img), img.shape[0], img.shape[1], C.byref(p), _fp(dst))
        return dst

    def octave_planes(self, first, params=None):
        """Returns (L[S+2,h,w], R[S+2,h,w], next[h//2,w//2])."""
        img = self._img(first)
        p = params or self.default_params()
        h, w = img.shape
        S = p.number_of_scales
        Lp = np.empty((S + 2, h, w), np.float32)
        Rp = np.empty((S + 2, h, w), np.float32)
        nx = np.empty((h // 2, w // 2), np.float32)
        self.lib.orc_octave_planes(_fp(img), h, w, C.byref(p), _fp(Lp), _fp(Rp), _fp(nx))
        return Lp, Rp, nx

    def find_affine_shape(self, blur, x, y, s, pd, params=None):
        img = self._img(blur)
        p = params or self.default_params()
        U = np.zeros(4, np.float32)
        it = C.c_int(0)
        ok = self.lib.orc_find_affine_shape(_fp(img), img.shape[0], img.shape[1], C.byref(p), x, y, s, pd, _fp(U), C.byref(it))
        return bool(ok), U, it.value

    def rectify(self, U):
        A = np.array(U, np.float32)
        self.lib.orc_rectify(_fp(A))
        return A

    def normalize_affine(self, image, x, y, s, A, params=None):
        img = self._img(image)
        p = params or self.default_params()
        patch = np.zeros((p.patch_size, p.patch_size), np.float32)
        rej = self.lib.orc_normalize_affine(_fp(img), img.shape[0], img.shape[1], C.byref(p), x, y, s,
                                            float(A[0]), float(A[1]), float(A[2]), float(A[3]), _fp(patch))
        return bool(rej), patch

    def sift(self, patch, params=None):
        p = params or self.default_params()
        pt = np.array(patch, np.float32, order="C")
        desc = np.zeros(128, np.uint8)
        self.lib.orc_sift(_fp(pt), C.byref(p), desc.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return desc, pt


def build(ref=None):
    """Compile the port (always) and the reference build (when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref is None:
        ref = os.path.isdir("/root/reference")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have(kind):
    return os.path.exists(_path(kind))


def _path(kind):
    return os.path.join(HERE, "libhesaff_oracle.so") if kind == "port" else os.path.join(HERE, "_ref", "libhesaff_ref.so")


_cache = {}


def load(kind="port"):
    if kind not in _cache:
        if not have(kind):
            build(ref=(kind == "ref"))
        _cache[kind] = Oracle(_path(kind))
    return _cache[kind]
