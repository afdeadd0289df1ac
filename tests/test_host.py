"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/hesaff_b200.h
declares, the parameter block mirrors the reference structs, the sharding logic works over gloo (world 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    import hesaff_b200
    hdr = open(os.path.join(ROOT, "include", "hesaff_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hesaff_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(hesaff_b200.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(hesaff_b200.EXPORTED_SYMBOLS) == declared
    assert lib.hesaff_abi_version() == 1


def test_blur_kernel_sass_uses_tma_and_packed_fp32():
    """Static proof on the built library (no GPU needed): every k_blur_tma instantiation stages its tile with one TMA
    box load (UTMALDG.3D) behind an mbarrier (SYNCS.*) and runs its filter passes on packed f32x2 math."""
    import shutil
    import hesaff_b200
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", hesaff_b200.lib_path()], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    blur = [f for f in funcs if f.startswith("_Z10k_blur_tma")]
    assert len(blur) >= 9, len(blur)                      # 1..21 taps
    for f in blur:
        assert f.count("UTMALDG.3D") == 1, f[:80]
        assert "SYNCS.ARRIVE.TRANS64" in f and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in f, f[:80]
    wide = [f for f in blur if "ILi11E" in f or "ILi15E" in f]
    assert wide and all("FFMA2" in f and "FADD2" in f for f in wide)
    # nothing on the path is a dense contraction: no tensor-core instructions anywhere in the library
    assert not re.search(r"\b(HMMA|IMMA|UTCHMMA|UTCMMA|QGMMA|HGMMA)\b", sass)


def test_header_is_plain_c():
    """The boundary is a C ABI: the header compiles as C99 on its own (no C++ or torch types in the signatures)."""
    r = subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Werror", "-fsyntax-only",
                        os.path.join(ROOT, "include", "hesaff_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_tools_do_not_import_the_oracle():
    """tools/ is product-side tooling (generators, profilers): the oracle stays behind tests/, smoke() and bench.py."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert "from oracle" not in src and "import oracle" not in src, f


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference's CPU code on host cores) on the smallest workload: exactly one stdout
    line, carrying the keys the contract names."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-images", "2", "--workload", "single_640x480"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("single_640x480")


def test_bench_roofline_traffic_comes_from_a_committed_ncu_capture():
    """roofline.traffic is read from the ncu --set full summary under profiles/ that bench.py names."""
    sys.path.insert(0, ROOT)
    import bench
    traffic, algorithmic = bench.ncu_traffic()
    assert traffic and algorithmic and 0.8 < traffic / algorithmic < 1.2, (traffic, algorithmic)
    blur, survey, n_o = bench.pyramid_algorithmic_bytes(1920, 1080, 3, 5, 0)
    assert len(n_o) == 7 and blur == 160342800 and survey == 214941000      # 214.9 MB per 1080p image


def test_params_default_mirror_reference_structs(port_oracle):
    import hesaff_b200
    p = hesaff_b200.HessianAffineParams()
    o = port_oracle.default_params()
    for name in ("threshold", "max_iter", "desc_factor", "patch_size", "number_of_scales", "initial_sigma",
                 "edge_eigenvalue_ratio", "border", "convergence_threshold", "smm_window_size", "max_octaves"):
        assert getattr(p, name) == getattr(o, name), name
    assert ctypes.sizeof(hesaff_b200.HessianAffineParams) == 48
    assert hesaff_b200.KEYPOINT_DTYPE.itemsize == 164      # struct Keypoint, hesaff.cpp:41-48


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import torch
    import hesaff_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hesaff_b200.HesaffError, match="no CUDA device|CUDA"):
        hesaff_b200.AffineHessianDetector()
    # the consumer entry point likewise: an error code, not a host computation
    import ctypes as C
    L = hesaff_b200.lib()
    L.hesaff_match_descriptors.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    buf = (C.c_char * 1024)()
    rc = L.hesaff_match_descriptors(0, buf, 1, buf, 1, buf, buf, None, None)
    assert rc < 0 and b"no CPU fallback" in L.hesaff_last_error()
    # and the host program: exit code 2 with the library's message (the reference has no such path; a silent empty
    # result would look like a valid run)
    import subprocess
    exe = os.path.join(ROOT, "hesaff_b200", "host", "hesaff")
    if os.path.exists(exe):
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "tex_320x240_s11.pgm")], capture_output=True, text=True, timeout=60)
        assert r.returncode == 2 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)
        out = os.path.join(ROOT, "tests", "golden", "tex_320x240_s11.pgm.hesaff.sift")
        assert not os.path.exists(out)


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may touch oracle/."""
    pkg = os.path.join(ROOT, "hesaff_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "from oracle" not in src and "import oracle" not in src and "oracle_api.h" not in src, f
                assert "libhesaff_oracle" not in src and "libhesaff_ref" not in src, f


def test_sift_file_writer_format(tmp_path):
    import hesaff_b200
    k = np.zeros(2, hesaff_b200.KEYPOINT_DTYPE)
    k["x"] = [12.5, 1234.5678]; k["y"] = [7.25, 0.001234567]; k["s"] = [2.0, 3.0]
    k["a11"] = [1.0, 1.25]; k["a21"] = [0.0, 0.1]; k["a22"] = [1.0, 0.8]
    k["desc"][0, :3] = [1, 2, 255]
    path = str(tmp_path / "t.sift")
    n = hesaff_b200.lib().hesaff_write_sift_file(path.encode(), k.ctypes.data, 2, ctypes.c_float(3.0 * np.sqrt(3.0)))
    assert n == 2
    lines = open(path).read().split("\n")
    assert lines[0] == "128" and lines[1] == "2" and lines[4] == ""
    t0 = lines[2].split()
    assert len(t0) == 133 and t0[0] == "12.5" and t0[1] == "7.25" and t0[5:8] == ["1", "2", "255"]
    assert lines[3].split()[0] == "1234.57" and lines[3].split()[1] == "0.00123457"   # ostream default: 6 significant digits
    sc2 = (3.0 * np.sqrt(3.0) * 2.0) ** 2
    assert abs(float(t0[2]) - 1.0 / sc2) < 1e-7 and float(t0[3]) == 0.0


def test_partition_and_offsets():
    from hesaff_b200 import shard
    assert shard.partition(8192, 8).tolist() == [1024 * i for i in range(9)]
    assert shard.partition(10, 4).tolist() == [0, 3, 6, 8, 10]
    assert shard.partition(3, 8).tolist() == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    c = np.array([[5, 3], [9, 9], [0, 0], [4, 1]])
    assert shard.global_offsets(c).tolist() == [0, 3, 12, 12, 13]


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hesaff_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
n_total = 7
starts = shard.partition(n_total, 2)
rng = np.random.default_rng(0)
all_counts = rng.integers(0, 1000, (n_total, 2)).astype(np.int32)
mine = all_counts[starts[rank]:starts[rank + 1]]
g = shard.all_gather_counts(dist, torch, mine, "cpu")
assert g.shape == (n_total, 2) and np.array_equal(g, all_counts), (rank, g)
off = shard.global_offsets(g)
assert off[-1] == all_counts[:, 1].sum() and off[starts[1]] == all_counts[:starts[1], 1].sum()
# the step after the path: variable-size all-gather of the Keypoint records (SURVEY 8(f) rank 3)
from hesaff_b200 import KEYPOINT_DTYPE
recs = np.zeros(int(off[-1]), KEYPOINT_DTYPE)
recs["x"] = np.arange(len(recs)); recs["type"] = 7; recs["desc"] = (np.arange(len(recs)) % 251)[:, None]
lo, hi = off[starts[rank]], off[starts[rank + 1]]
got = shard.all_gather_keypoints(dist, torch, recs[lo:hi], g, starts, "cpu")
assert tuple(got.shape) == (len(recs), 164)
back = got.numpy().reshape(-1).view(KEYPOINT_DTYPE)
assert back.tobytes() == recs.tobytes(), rank
try:
    shard.all_gather_keypoints(dist, torch, recs[lo:hi][:-1], g, starts, "cpu")
    raise SystemExit("size mismatch not detected")
except ValueError:
    pass
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_count_all_gather_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_pnm_header_parser():
    """hesaff_pnm_info is host-only (no GPU): P5/P6, comments, whitespace forms, and the error cases."""
    import ctypes as C
    import hesaff_b200
    L = hesaff_b200.lib()
    L.hesaff_pnm_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]

    def info(b):
        w, h, ch, off = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        rc = L.hesaff_pnm_info(b, len(b), C.byref(w), C.byref(h), C.byref(ch), C.byref(off))
        return rc, w.value, h.value, ch.value, off.value

    hdr = b"P5\n# made by a test\n  7 3\n#x\n255\n"
    assert info(hdr + bytes(21)) == (0, 7, 3, 1, len(hdr))
    assert info(b"P6 4 2 255 " + bytes(24)) == (0, 4, 2, 3, 11)
    assert info(b"P6 4 2 255 " + bytes(23))[0] < 0            # truncated payload
    assert info(b"P2 4 2 255 " + bytes(24))[0] < 0            # ASCII PNM is not supported
    assert info(b"P5 4 2 65535 " + bytes(16))[0] < 0          # 16-bit
    assert info(b"P5 4")[0] < 0 and info(b"")[0] < 0
    assert b"PNM" in L.hesaff_last_error()
