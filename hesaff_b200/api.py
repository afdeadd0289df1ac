"""Host-side mirror of the reference's interface for the detect -> affine -> describe path, bound to the
C-ABI shared library (include/hesaff_b200.h) with ctypes.  Names follow hesaff.cpp:

    HessianAffineParams            hesaff.cpp:21-36  (+ the struct defaults the CLI never changes)
    AffineHessianDetector          hesaff.cpp:50-131 : detectPyramidKeypoints(), keys, exportKeypoints()

There is no CPU fallback: importing works anywhere (so that the symbols can be checked), but creating a
detector raises HesaffError unless the CUDA library is present and a B200-class device is visible.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

KEYPOINT_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("s", "f4"), ("a11", "f4"), ("a12", "f4"), ("a21", "f4"),
                           ("a22", "f4"), ("response", "f4"), ("type", "i4"), ("desc", "u1", (128,))])
DETECTION_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("s", "f4"), ("pd", "f4"), ("type", "i4"), ("response", "f4"),
                            ("affine_ok", "i4"), ("u11", "f4"), ("u12", "f4"), ("u21", "f4"), ("u22", "f4"),
                            ("iters", "i4"), ("described", "i4"), ("a11", "f4"), ("a12", "f4"), ("a21", "f4"),
                            ("a22", "f4"), ("desc", "u1", (128,))])
assert KEYPOINT_DTYPE.itemsize == 164 and DETECTION_DTYPE.itemsize == 196

HESSIAN_DARK, HESSIAN_BRIGHT, HESSIAN_SADDLE = 0, 1, 2   # pyramid.h:51-55

EXPORTED_SYMBOLS = [
    "hesaff_abi_version", "hesaff_last_error", "hesaff_params_default", "hesaff_create", "hesaff_destroy",
    "hesaff_detect_u8", "hesaff_detect_f32", "hesaff_detect_rgb8", "hesaff_pnm_info", "hesaff_detect_pnm", "hesaff_match_descriptors", "hesaff_result_counts", "hesaff_result_total", "hesaff_result_keypoints",
    "hesaff_result_keypoints_device", "hesaff_set_host_output", "hesaff_result_ellipses", "hesaff_result_detections", "hesaff_debug_geometry",
    "hesaff_debug_octave_size", "hesaff_debug_plane", "hesaff_debug_patches", "hesaff_launch_count",
    "hesaff_set_profiling", "hesaff_stage_times_ms", "hesaff_blur_time_ms", "hesaff_write_sift_file",
    "hesaff_result_sift_text", "hesaff_export_sift_file", "hesaff_write_keypoints_binary", "hesaff_read_keypoints_binary",
    "hesaff_debug_format_floats",
]


class HesaffError(RuntimeError):
    pass


class HessianAffineParams(C.Structure):
    """hesaff_params: HessianAffineParams (hesaff.cpp:21-36) + PyramidParams / AffineShapeParams defaults."""
    _fields_ = [("threshold", C.c_float), ("max_iter", C.c_int), ("desc_factor", C.c_float), ("patch_size", C.c_int),
                ("verbose", C.c_int), ("number_of_scales", C.c_int), ("initial_sigma", C.c_float),
                ("edge_eigenvalue_ratio", C.c_float), ("border", C.c_int), ("convergence_threshold", C.c_float),
                ("smm_window_size", C.c_int), ("max_octaves", C.c_int)]

    def __init__(self, **kw):
        super().__init__()
        lib().hesaff_params_default(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


_lib = None


def lib_path():
    return _build.LIB


def lib():
    """Loads the C-ABI library; raises HesaffError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB):
            raise HesaffError("CUDA library %s is missing: run `python -m hesaff_b200.build` (there is no CPU fallback)" % _build.LIB)
        L = C.CDLL(_build.LIB)
        L.hesaff_last_error.restype = C.c_char_p
        L.hesaff_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(HessianAffineParams)] + [C.c_int] * 5
        L.hesaff_destroy.argtypes = [C.c_void_p]
        for name in ("hesaff_detect_u8", "hesaff_detect_f32", "hesaff_detect_rgb8"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
        L.hesaff_result_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hesaff_result_total.argtypes = [C.c_void_p]
        L.hesaff_result_total.restype = C.c_int64
        L.hesaff_result_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.hesaff_result_keypoints_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.hesaff_set_host_output.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.hesaff_result_ellipses.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.hesaff_result_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64)]
        L.hesaff_debug_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hesaff_debug_octave_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hesaff_debug_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hesaff_debug_patches.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.hesaff_launch_count.argtypes = [C.c_void_p, C.c_int]
        L.hesaff_launch_count.restype = C.c_int64
        L.hesaff_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.hesaff_stage_times_ms.argtypes = [C.c_void_p, C.c_void_p]
        L.hesaff_blur_time_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.hesaff_write_sift_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_float]
        L.hesaff_result_sift_text.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.hesaff_export_sift_file.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.hesaff_write_keypoints_binary.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        L.hesaff_read_keypoints_binary.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.hesaff_debug_format_floats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise HesaffError("hesaff_b200 error %d: %s" % (rc, lib().hesaff_last_error().decode()))
    return rc


class AffineHessianDetector:
    """Batch counterpart of AffineHessianDetector (hesaff.cpp:50-131).

    detectPyramidKeypoints(images) runs detect -> affine shape -> patch normalisation -> SIFT for a batch of
    gray images (numpy u8/f32 arrays [N,H,W] or [H,W], or a torch CUDA tensor) and fills `counts`;
    `keys()` returns the Keypoint records (hesaff.cpp:41-48) of all images, reference order.
    """

    def __init__(self, params=None, device=0, max_width=1920, max_height=1080, max_batch=0, max_candidates_per_image=0):
        self.par = params or HessianAffineParams()
        self._h = C.c_void_p()
        rc = lib().hesaff_create(C.byref(self._h), C.byref(self.par), device, max_width, max_height, max_batch,
                                 max_candidates_per_image)
        if rc < 0:
            msg = lib().hesaff_last_error().decode()
            if self._h:
                lib().hesaff_destroy(self._h)
                self._h = C.c_void_p()
            raise HesaffError("hesaff_create failed (%d): %s" % (rc, msg))
        self.n_images = 0
        self.n_detected = np.zeros(0, np.int32)    # g_numberOfPoints per image, hesaff.cpp:68
        self.n_described = np.zeros(0, np.int32)   # g_numberOfAffinePoints per image, hesaff.cpp:103

    def close(self):
        if getattr(self, "_h", None):
            lib().hesaff_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- the hot path ----
    def detectPyramidKeypoints(self, images, stream=None):
        on_device = 0
        if hasattr(images, "data_ptr"):   # torch tensor
            import torch
            t = images
            rgb = t.dim() == 4 and t.shape[3] == 3 and t.dtype == torch.uint8   # [N,H,W,3] interleaved colour
            if t.dim() == 2:
                t = t.unsqueeze(0)
            if not (t.dim() == 3 or rgb) or t.stride(-1) != 1:
                raise HesaffError("expected a [H,W], [N,H,W] or uint8 [N,H,W,3] tensor with unit innermost stride")
            if rgb and t.stride(2) != 3:
                raise HesaffError("interleaved colour input needs a pixel stride of 3 bytes")
            if t.dtype == torch.uint8:
                fn = lib().hesaff_detect_rgb8 if rgb else lib().hesaff_detect_u8
            elif t.dtype == torch.float32:
                fn = lib().hesaff_detect_f32
            else:
                raise HesaffError("unsupported tensor dtype %s (uint8 gray / uint8 RGB / float32 gray)" % t.dtype)
            on_device = 1 if t.is_cuda else 0
            n, h, w = t.shape[:3]
            esz = t.element_size()
            ptr, rp, ist = t.data_ptr(), t.stride(1) * esz, t.stride(0) * esz
            if n == 1:
                ist = max(ist, rp * h)
            if on_device and stream is None:
                # order the call after the work already queued on the tensor's current torch stream; torch's default
                # stream is the legacy default stream, whose handle 0 the C-ABI reads as "the context's own stream"
                stream = torch.cuda.current_stream(t.device).cuda_stream or 1    # 1 = cudaStreamLegacy
            self._keep = t
        else:
            a = np.asarray(images)
            if a.ndim == 2:
                a = a[None]
            rgb = a.ndim == 4 and a.shape[3] == 3 and a.dtype == np.uint8   # [N,H,W,3] interleaved colour: gray on the GPU
            assert a.ndim == 3 or rgb
            if a.dtype != np.uint8:
                a = a.astype(np.float32, copy=False)
            a = np.ascontiguousarray(a)
            n, h, w = a.shape[:3]
            fn = lib().hesaff_detect_rgb8 if rgb else (lib().hesaff_detect_u8 if a.dtype == np.uint8 else lib().hesaff_detect_f32)
            rp = w * a.itemsize * (3 if rgb else 1)
            ptr, ist = a.ctypes.data, rp * h   # C-contiguous (a[None] reports a zero stride for the new axis)
            self._keep = a
        _check(fn(self._h, C.c_void_p(ptr), n, w, h, rp, ist, on_device, C.c_void_p(stream or 0)))
        self.n_images = n
        self.n_detected = np.zeros(n, np.int32)
        self.n_described = np.zeros(n, np.int32)
        _check(lib().hesaff_result_counts(self._h, self.n_detected.ctypes.data, self.n_described.ctypes.data))
        return self

    def detectFiles(self, files, stream=None):
        """`files`: list of bytes objects (or paths) holding binary PNM files of one size -- the pixel payload goes to
        the GPU as it lies in the file, the gray conversion of hesaff.cpp:137-148 runs there (hesaff_detect_pnm)."""
        blobs = [open(f, "rb").read() if isinstance(f, str) else bytes(f) for f in files]
        n = len(blobs)
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in blobs])
        sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
        self._keep = blobs
        _check(lib().hesaff_detect_pnm(self._h, ptrs, sizes, n, C.c_void_p(stream or 0)))
        self.n_images = n
        self.n_detected = np.zeros(n, np.int32)
        self.n_described = np.zeros(n, np.int32)
        _check(lib().hesaff_result_counts(self._h, self.n_detected.ctypes.data, self.n_described.ctypes.data))
        return self

    def total(self):
        return int(_check(lib().hesaff_result_total(self._h)))

    def keys(self, out=None):
        """All Keypoint records (KEYPOINT_DTYPE), image-major; image i = keys[offsets[i]:offsets[i+1]]."""
        n = self.total()
        if out is None:
            out = np.empty(n, KEYPOINT_DTYPE)
        _check(lib().hesaff_result_keypoints(self._h, out.ctypes.data, len(out)))
        return out[:n]

    def set_host_output(self, out):
        """Stream every chunk's records into `out` (KEYPOINT_DTYPE array, ideally pinned) during detect; None disables."""
        self._host_out = out
        if out is None:
            _check(lib().hesaff_set_host_output(self._h, None, 0))
        else:
            _check(lib().hesaff_set_host_output(self._h, out.ctypes.data, len(out)))

    def keys_device_ptr(self):
        p = C.c_void_p()
        _check(lib().hesaff_result_keypoints_device(self._h, C.byref(p)))
        return p.value

    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.n_described)]).astype(np.int64)

    def ellipses(self):
        """(u, v, a, b, c) per keypoint, as exportKeypoints computes them (hesaff.cpp:115-125)."""
        n = self.total()
        out = np.empty((n, 5), np.float32)
        _check(lib().hesaff_result_ellipses(self._h, out.ctypes.data, n))
        return out

    def detections(self):
        """Every detection with its per-stage results (DETECTION_DTYPE), for parity tests."""
        nt = C.c_int64()
        _check(lib().hesaff_result_detections(self._h, None, 0, C.byref(nt)))
        out = np.empty(nt.value, DETECTION_DTYPE)
        _check(lib().hesaff_result_detections(self._h, out.ctypes.data, len(out), C.byref(nt)))
        return out

    def exportKeypoints(self, path, image=0, on_host=False):
        """Writes <path> in the reference's .hesaff.sift format (hesaff.cpp:107-130) for one image.  The text is
        formatted on the GPU (hesaff_export_sift_file); on_host=True uses the host ostream writer instead (same bytes)."""
        if not on_host:
            return _check(lib().hesaff_export_sift_file(self._h, image, path.encode()))
        k = self.keys()
        o = self.offsets()
        k = np.ascontiguousarray(k[o[image]:o[image + 1]])
        return _check(lib().hesaff_write_sift_file(path.encode(), k.ctypes.data, len(k), self.par.desc_factor))

    def siftText(self, image=0):
        """The .hesaff.sift file content of one image as bytes, formatted on the GPU."""
        nb = C.c_size_t()
        _check(lib().hesaff_result_sift_text(self._h, image, None, 0, C.byref(nb)))
        buf = C.create_string_buffer(nb.value)
        _check(lib().hesaff_result_sift_text(self._h, image, buf, nb.value, C.byref(nb)))
        return buf.raw[:nb.value]

    def exportKeypointsBinary(self, path, image=0):
        """Binary sidecar (hesaff_write_keypoints_binary): header + the 164-byte Keypoint records of one image."""
        k = self.keys()
        o = self.offsets()
        k = np.ascontiguousarray(k[o[image]:o[image + 1]])
        return _check(lib().hesaff_write_keypoints_binary(path.encode(), k.ctypes.data, len(k)))

    def formatFloats(self, values):
        """Diagnostic: the device "%g" formatter on an array of float32 (list of str)."""
        v = np.ascontiguousarray(values, np.float32)
        out = np.zeros((len(v), 16), np.uint8)
        _check(lib().hesaff_debug_format_floats(self._h, v.ctypes.data, len(v), out.ctypes.data))
        return [bytes(r).split(b"\0")[0].decode() for r in out]

    # ---- stage access (tests) ----
    def geometry(self):
        no, nl = C.c_int(), C.c_int()
        _check(lib().hesaff_debug_geometry(self._h, C.byref(no), C.byref(nl)))
        sizes = []
        for o in range(no.value):
            w, h = C.c_int(), C.c_int()
            _check(lib().hesaff_debug_octave_size(self._h, o, C.byref(w), C.byref(h)))
            sizes.append((h.value, w.value))
        return no.value, nl.value, sizes

    def plane(self, image, octave, level, kind):
        _, _, sizes = self.geometry()
        out = np.empty(sizes[octave], np.float32)
        _check(lib().hesaff_debug_plane(self._h, image, octave, level, {"L": 0, "R": 1}[kind], out.ctypes.data))
        return out

    def patches(self, normalized=False):
        n = self.total()
        out = np.empty((n, 41, 41), np.float32)
        _check(lib().hesaff_debug_patches(self._h, int(normalized), out.ctypes.data, n))
        return out

    def launch_count(self, reset=False):
        return int(lib().hesaff_launch_count(self._h, int(reset)))

    def set_profiling(self, on):
        _check(lib().hesaff_set_profiling(self._h, int(on)))

    def blur_time_ms(self):
        ms, n = C.c_float(), C.c_int()
        _check(lib().hesaff_blur_time_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stage_times_ms(self):
        out = np.zeros(6, np.float32)
        _check(lib().hesaff_stage_times_ms(self._h, out.ctypes.data))
        return out
