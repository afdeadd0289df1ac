import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hesaff_b200 as hb
from tools.gen_textured import textured
img = textured(333, 251, 7)
det = hb.AffineHessianDetector(hb.HessianAffineParams(), 0, 333, 251, 1)
det.detectPyramidKeypoints(img)
print("ok", det.n_detected, det.n_described)
