"""hesaff_b200: B200-native Hessian-Affine + SIFT hot path (detect -> affine -> describe) behind a C-ABI.

The product is hesaff_b200/libhesaff_b200.so (hand-written sm_100a CUDA, csrc/) declared in
include/hesaff_b200.h; this package is the thin host-side mirror of the reference's interface.
"""
from .api import (AffineHessianDetector, HessianAffineParams, HesaffError, KEYPOINT_DTYPE, DETECTION_DTYPE,  # noqa: F401
                  EXPORTED_SYMBOLS, lib, lib_path)
