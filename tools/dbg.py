import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hesaff_b200 as hb
from oracle import oracle
from tools.gen_textured import textured
from parity import compare_keypoints
port = oracle.load("port")
for (w,h,seed,over) in [(40,30,25,{}),(320,240,11,{}),(640,480,1,{}),(1920,1080,2,{})]:
    img = textured(w, h, seed)
    det = hb.AffineHessianDetector(hb.HessianAffineParams(**over), 0, w, h, 1)
    det.set_profiling(True)
    det.detectPyramidKeypoints(img)
    got = det.detections(); want = port.detect(img.astype(np.float32))
    print(w,h,"det",len(got),len(want),"affine",got["affine_ok"].sum(),want["affine_ok"].sum(),"desc",got["described"].sum(),want["described"].sum())
    n=min(len(got),len(want))
    print("  affine flips", (got["affine_ok"][:n]!=want["affine_ok"][:n]).sum(), "iters diff", (got["iters"][:n]!=want["iters"][:n]).sum(), "desc flips", (got["described"][:n]!=want["described"][:n]).sum())
    both=(got["affine_ok"][:n]==1)&(want["affine_ok"][:n]==1)
    print("  max dU", max(np.abs(got[f][:n][both]-want[f][:n][both]).max() for f in ("u11","u12","u21","u22")), "s equal", np.array_equal(got["s"][:n],want["s"][:n]), np.abs(got["s"][:n]-want["s"][:n]).max())
    print("  ", compare_keypoints(det.keys(), want[want["described"]==1]))
    print("  stage ms", det.stage_times_ms(), "launches", det.launch_count())
    det.close()
