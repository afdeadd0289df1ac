#!/usr/bin/env python
"""bench.py -- throughput of the detect -> affine -> describe hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic textured images per GPU (weak scaling:
every rank processes its own `batch` images; images are independent, so there is no data-path collective,
only the all-gather of per-image keypoint counts).  Prints ONE JSON line on rank 0.

  value : whole-job Mpixels/s with the u8 images already resident in HBM (device-timed, max over ranks)
  e2e   : the same metric through the C-ABI with HOST buffers: pinned host images -> H2D -> path -> D2H of
          every Keypoint record, all inside the timed region
  roofline      : pyramid blur+response kernels (k_blur), algorithmic HBM bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  : the reference's own CPU code (oracle/_ref, else the plain-C port) on this box's host cores, bounded sample
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, images per GPU, param overrides)           BASELINE.json configs[...]
    "batch1024_1080p": (1920, 1080, 1024, {}),                                              # [2] (and [3] at 8 GPUs)
    "single_4k_S10_oct3": (3840, 2160, 1, {"number_of_scales": 10, "max_octaves": 3}),     # [1]
    "dense_4096_thr5_oct6": (4096, 4096, 8, {"threshold": 5.0, "max_octaves": 6}),         # [4] (batch of 8)
    "single_640x480": (640, 480, 1, {}),                                                    # [0]
}


def synth_textured_gpu(torch, n, h, w, seed, device, chunk=32):
    """Textured images like tools/gen_textured.textured (unit-variance band-limited noise at sigma 1..32 px,
    mean 128, std 48, clipped to u8), generated on the device in the frequency domain (circular boundary)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    fy = torch.fft.fftfreq(h, device=device).view(h, 1)
    fx = torch.fft.rfftfreq(w, device=device).view(1, w // 2 + 1)
    f2 = fx * fx + fy * fy
    out = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        acc = torch.zeros((m, h, w), device=device)
        for sigma in (1, 2, 4, 8, 16, 32):
            noise = torch.randn((m, h, w), generator=g, device=device)
            comp = torch.fft.irfft2(torch.fft.rfft2(noise) * torch.exp(-2.0 * (3.141592653589793 * sigma) ** 2 * f2), s=(h, w))
            acc += comp / comp.std(dim=(1, 2), keepdim=True)
            del noise, comp
        z = (acc - acc.mean(dim=(1, 2), keepdim=True)) / acc.std(dim=(1, 2), keepdim=True)
        out[s:s + m] = (128 + 48 * z).clamp_(0, 255).to(torch.uint8)
        del acc, z
    return out


def pyramid_algorithmic_bytes(w, h, S, border, max_octaves):
    """SURVEY.md 8(d), minimal-materialisation model, per image.
    blur kernels: u8 read + float image (1+4) on octave 0... here: every blur launch reads one fp32 plane (4 B/px) and
    writes L (4) and R (4); the first blur reads the fp32 gray image; the level-S launch also writes the decimated
    seed (N_o/4 * 4 B); R[0] of octaves >= 1 is one extra read + write of the small plane."""
    n_o, r, c, o = [], h, w, 0
    while r > 2 * border + 2 and c > 2 * border + 2 and (max_octaves <= 0 or o < max_octaves):
        n_o.append(r * c)
        r //= 2
        c //= 2
        o += 1
    blur = 12 * n_o[0] if n_o else 0                       # first blur: read 4, write L 4 + R 4
    for i, n in enumerate(n_o):
        blur += (S + 1) * 12 * n                           # S+1 incremental blurs
        if i + 1 < len(n_o):
            blur += n_o[i + 1] * 4                         # decimated seed written by the level-S launch
    survey = 5 * (n_o[0] if n_o else 0) + sum((16 * S + 26) * n for n in n_o)   # incl. u8 ingest and the NMS reads
    return blur, survey, n_o


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, power) if p > 250] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------
# CPU side: the reference's own code on host cores (test/measurement infrastructure, never the product path)
# --------------------------------------------------------------------------------------------------------
_CV2_CB = None


def _install_cv2_blur(orc):
    """--cpu-blur cv2: the reference build calls the real OpenCV's GaussianBlur (single-threaded) instead of the stand-in's
    (oracle/shim/cv_shim.cpp: orc_shim_set_blur).  BASELINE.md 3.1; the ctypes round trip adds a few microseconds per call."""
    global _CV2_CB
    if _CV2_CB is not None:
        return
    import ctypes as C
    import cv2
    import numpy as np
    cv2.setNumThreads(1)
    FP = C.POINTER(C.c_float)

    def blur(src, dst, rows, cols, sstep, dstep, ksize, sigma):
        a = np.ctypeslib.as_array(src, shape=(rows, sstep // 4))[:, :cols]
        d = np.ctypeslib.as_array(dst, shape=(rows, dstep // 4))[:, :cols]
        cv2.GaussianBlur(a, (ksize, ksize), sigma, dst=d, sigmaY=sigma, borderType=cv2.BORDER_REPLICATE)

    _CV2_CB = C.CFUNCTYPE(None, FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double)(blur)
    orc.lib.orc_shim_set_blur(_CV2_CB)


def _cpu_worker(args):
    kind, img_bytes, h, w, over = args
    import numpy as np
    from oracle import oracle
    orc = oracle.load(kind)
    if os.environ.get("HESAFF_CPU_BLUR") == "cv2" and kind == "ref":
        _install_cv2_blur(orc)
    img = np.frombuffer(img_bytes, np.uint8).reshape(h, w).astype(np.float32)
    t = time.perf_counter()
    d = orc.detect(img, orc.default_params(**over))
    return time.perf_counter() - t, int(len(d)), int(d["described"].sum())


NCU_TRAFFIC_FILE = "profiles/r2_k_blur_tma_ncu.txt"     # first launch listed = octave 0, 11 taps, 32 x 1920x1080


def ncu_traffic():
    """DRAM bytes (read+write) of one octave-0 k_blur_tma launch from the committed `ncu --set full` capture
    (NCU_TRAFFIC_FILE), with the algorithmic bytes of the same launch."""
    path = os.path.join(ROOT, NCU_TRAFFIC_FILE)
    try:
        rd = wr = None
        for line in open(path):
            t = line.split()
            if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(t[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[t[2]]
                if t[0].startswith("dram__bytes_read") and rd is None:
                    rd = v
                elif t[0].startswith("dram__bytes_write") and wr is None:
                    wr = v
            if rd is not None and wr is not None:
                return rd + wr, 32 * 1920 * 1080 * 12.0
    except (OSError, ValueError, KeyError):
        pass
    return None, None


def physical_cores():
    """Worker count for the CPU reference: one process per physical core (SMT siblings only add contention for
    this memory-bound code; measured here: 64 processes beat 128 on a 64-core/128-thread host)."""
    n = os.cpu_count() or 1
    try:
        out = subprocess.run(["lscpu", "-p=CORE,SOCKET"], capture_output=True, text=True, timeout=10).stdout
        phys = len({ln for ln in out.splitlines() if ln and not ln.startswith("#")})
        if 0 < phys <= n:
            return phys
    except (OSError, subprocess.SubprocessError):
        pass
    return n


def cpu_reference_run(images_u8, over, cores):
    """Runs the reference CPU path over `images_u8` ([n,h,w] numpy) with `cores` worker processes (the
    reference is single-threaded: one process per image, as many at a time as there are cores)."""
    import multiprocessing as mp
    from oracle import oracle
    kind = "ref" if oracle.have("ref") else "port"
    oracle.load(kind)
    n, h, w = images_u8.shape
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(min(cores, n)) as pool:
        res = pool.map(_cpu_worker, [(kind, images_u8[i].tobytes(), h, w, over) for i in range(n)],
                       chunksize=max(1, n // min(cores, n)))
    wall = time.perf_counter() - t0
    return {"kind": "reference" if kind == "ref" else "port", "wall_s": wall, "images": n, "mpix": n * h * w / 1e6,
            "single_core_s_per_image": statistics.mean(r[0] for r in res), "detections": [r[1] for r in res],
            "described": [r[2] for r in res], "workers": min(cores, n)}


def main():
    # Exactly ONE line goes to stdout (the JSON record): library chatter on fd 1 (e.g. NCCL's version banner) is sent to
    # stderr for the duration of the run.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch1024_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override images per GPU")
    ap.add_argument("--cpu-images", type=int, default=0, help="CPU sample size (default: 16 images per worker process for the cpu_baseline leg, BASELINE.md 3.3; 8 per timed step of --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-blur", default="shim", choices=["shim", "cv2"],
                    help="--impl reference only: blur of the CPU build = the stand-in pinned bit-identical to OpenCV (default) or the real cv2")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W, H, batch, over = WORKLOADS[args.workload]
    if args.batch > 0:
        batch = args.batch
    S = over.get("number_of_scales", 3)
    host_cores = physical_cores()
    host_threads = os.cpu_count() or 1
    config = {"workload": "%s: %d x %dx%d u8 gray synthetic textured images per GPU, params %s" %
                          (args.workload, batch, W, H, over or "default"),
              "images_per_gpu": batch, "width": W, "height": H, "params": over,
              "parallelism": "image-parallel x%d (no data-path collective; all-gather of per-image counts)" % world,
              "l2": "inputs (%.0f MB/GPU) and pyramid planes exceed the 126 MB L2 every step" % (batch * W * H / 1e6)}

    import numpy as np

    # ---------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        # the reference's own CPU implementation on this box's host cores; rank 0 only
        if rank != 0:
            return 0
        if args.cpu_blur == "cv2":
            os.environ["HESAFF_CPU_BLUR"] = "cv2"
            config["cpu_blur"] = "cv2.GaussianBlur (OpenCV %s, 1 thread) through a ctypes callback" % __import__("cv2").__version__
        import torch
        from tools.gen_textured import textured
        # every timed step: 8 images per worker process (about 30 s of host work at 1080p); warm-up steps: one per process
        per_proc = 8 if W * H <= 1920 * 1080 else (2 if W * H <= 3840 * 2160 else 1)
        n_cpu = args.cpu_images or min(per_proc * host_cores, 512)
        if torch.cuda.is_available():
            imgs = synth_textured_gpu(torch, n_cpu, H, W, 1234, "cuda:0").cpu().numpy()
        else:
            imgs = np.stack([textured(W, H, 1000 + i) for i in range(min(n_cpu, 8))])
        times = []
        for it in range(args.warmup + args.steps):
            if it < args.warmup:
                cpu_reference_run(imgs[:min(host_cores, imgs.shape[0])], over, host_cores)
                continue
            r = cpu_reference_run(imgs, over, host_cores)
            times.append(r["wall_s"])
        t = sum(times) / len(times)
        v = imgs.shape[0] * H * W / 1e6 / t
        sample = "%d images of the workload per timed step, %d worker processes (one per physical core), %.1f images per process" % (
            imgs.shape[0], r["workers"], imgs.shape[0] / r["workers"])
        print(json.dumps({
            "impl": "reference", "metric": "Mpixels/sec end-to-end (detect+affine+SIFT)", "value": v, "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": r["workers"], "kind": r["kind"], "sample": sample,
                             "single_core_mpix_s": H * W / 1e6 / r["single_core_s_per_image"]},
            "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "keypoints_per_s": sum(r["described"]) / t,
        }))
        return 0

    # ---------------------------------------------------------------------------------------------------
    import torch
    import hesaff_b200 as hb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))

    par = hb.HessianAffineParams(**over)
    det = hb.AffineHessianDetector(par, device=local_rank, max_width=W, max_height=H, max_batch=0)
    images = synth_textured_gpu(torch, batch, H, W, 1234 + 7919 * rank, dev)
    host_in = torch.empty((batch, H, W), dtype=torch.uint8).pin_memory()
    host_in.copy_(images)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    counts_dev = torch.zeros((batch, 2), dtype=torch.int32, device=dev)
    gathered = [torch.zeros_like(counts_dev) for _ in range(world)] if world > 1 else None

    def step_device():
        det.detectPyramidKeypoints(images, stream=stream)
        if world > 1:   # the only collective of the path: per-image {detected, described} counts
            counts_dev.copy_(torch.from_numpy(np.stack([det.n_detected, det.n_described], 1)))
            dist.all_gather(gathered, counts_dev)

    host_out = None

    def step_e2e():
        # pinned host images -> (chunked, overlapped) H2D -> path -> every Keypoint record streamed D2H into pinned memory
        det.detectPyramidKeypoints(host_in, stream=stream)
        n = det.total()
        det.keys(out=host_out)          # already streamed during the call when host_out is registered
        if world > 1:
            counts_dev.copy_(torch.from_numpy(np.stack([det.n_detected, det.n_described], 1)))
            dist.all_gather(gathered, counts_dev)
        return n

    def timed(fn, steps, profile=False):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        det.set_profiling(profile)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stage = np.zeros(8, np.float64)
        for _ in range(steps):
            fn()
            if profile:
                stage[:6] += det.stage_times_ms()
                stage[6:8] += det.blur_time_ms()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        det.set_profiling(False)
        return ms, stage

    for _ in range(max(args.warmup, 0)):
        step_device()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    det.launch_count(reset=True)
    ms_dev, _ = timed(step_device, args.steps)
    launches = det.launch_count()
    clk = clocks.stop() if rank == 0 else None
    n_det, n_desc = int(det.n_detected.sum()), int(det.n_described.sum())
    # stage times / roofline: same steps again with per-stage CUDA events; profiling serialises the two chunk lanes so
    # that a stage's time is not inflated by the other lane's kernels
    ms_prof, stage = timed(step_device, args.steps, profile=True)

    n_first = det.total()
    t = torch.empty((int(n_first * 1.05) + 4096, hb.KEYPOINT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    host_out = t.numpy().view(hb.KEYPOINT_DTYPE).reshape(-1)
    det.set_host_output(host_out)
    step_e2e()   # warm
    ms_e2e, _ = timed(step_e2e, args.steps)
    n_e2e = det.total()
    det.set_host_output(None)

    mpix_step = batch * world * W * H / 1e6
    value = mpix_step / (ms_dev / args.steps / 1e3)
    e2e_v = mpix_step / (ms_e2e / args.steps / 1e3)

    if world > 1:
        tot = torch.tensor([n_det, n_desc], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        n_det, n_desc = int(tot[0]), int(tot[1])

    out = None
    if rank == 0:
        blur_bytes, survey_bytes, n_o = pyramid_algorithmic_bytes(W, H, S, par.border, par.max_octaves)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # per-launch figure: CUDA events around every k_blur launch of the profiled steps (29 per chunk at 1080p, S=3)
        blur_ms = stage[6] / args.steps
        n_blur_launches = int(round(stage[7] / args.steps))
        achieved = blur_bytes * batch / (blur_ms / 1e3) / 1e9 if blur_ms > 0 else 0.0
        pyr_nms_ms = float(stage[0] + stage[1] + stage[2]) / args.steps          # upload/convert + pyramid + NMS/localise
        survey_achieved = survey_bytes * batch / (pyr_nms_ms / 1e3) / 1e9 if pyr_nms_ms > 0 else 0.0
        out = {
            "metric": "Mpixels/sec end-to-end (detect+affine+SIFT)", "value": value, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_v, "unit": "Mpix/s", "h2d_bytes_per_step": batch * W * H,
                    "d2h_bytes_per_step": int(n_e2e) * hb.KEYPOINT_DTYPE.itemsize + batch * 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "keypoints_per_s": n_desc / (ms_dev / args.steps / 1e3),
            "detections_per_step": n_det, "described_per_step": n_desc,
            "ms_per_step_serialised_with_stage_events": ms_prof / args.steps,
            "stages_ms_per_step": {k: float(v) / args.steps for k, v in
                                   zip(("upload_convert", "pyramid", "nms_localize", "affine", "patch_sift", "compact"), stage[:6])},
            "roofline": {"kernel": "k_blur_tma<N> (TMA-staged separable Gaussian + det-Hessian epilogue): all %d launches of a step, "
                                   "sum of algorithmic bytes / sum of per-launch CUDA-event durations" % n_blur_launches,
                         "avg_launch_us": 1e3 * blur_ms / max(1, n_blur_launches), "algorithmic_bytes_per_launch_avg": blur_bytes * batch / max(1, n_blur_launches),
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_image": blur_bytes, "survey_bytes_per_image_incl_nms": survey_bytes,
                         "traffic": ncu_traffic()[0],
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE octave-0 k_blur_tma<11> launch over "
                                         "32 x 1920x1080 (ncu --set full, %s); algorithmic bytes of that launch: %.0f" % (
                                             NCU_TRAFFIC_FILE, ncu_traffic()[1] or 0),
                         # the same stage under SURVEY.md 8(d)'s wider definition: B_pyr (u8 ingest + every plane written once
                         # and read once per consumer, NMS reads included) over the pyramid + NMS/localisation stage times
                         "survey_b_pyr": {"bytes_per_image": survey_bytes, "stage_ms": pyr_nms_ms,
                                          "achieved": survey_achieved, "frac": survey_achieved / peak if peak else None}},
            "clocks": clk,
            "host_cores": host_cores, "host_threads": host_threads,
        }
        if not args.no_cpu_baseline:
            # >= 16 images per process at N=1 (about a minute of host work); at N>1 the baseline is only
            # the per-rank counts check (one image per process), the CPU figure belongs to the N=1 line
            per_proc = (16 if W * H <= 1920 * 1080 else (2 if W * H <= 3840 * 2160 else 1)) if world == 1 else 1
            n_cpu = min(args.cpu_images or per_proc * host_cores, batch)
            sample = images[:n_cpu].cpu().numpy()
            r = cpu_reference_run(sample, over, host_cores)
            cpu_v = r["mpix"] / r["wall_s"]
            gpu_det = det.n_detected[:n_cpu].tolist()        # rank 0's own images (the last call ran on the same batch)
            out["cpu_baseline"] = {
                "value": cpu_v, "unit": "Mpix/s", "cores": r["workers"], "kind": r["kind"],
                "sample": "first %d images of rank 0's batch over %d worker processes (%.1f per process), wall %.1f s" % (
                    n_cpu, r["workers"], n_cpu / r["workers"], r["wall_s"]),
                "single_core_mpix_s": H * W / 1e6 / r["single_core_s_per_image"],
                "counts_match_gpu": gpu_det == r["detections"],
            }
        print(json.dumps(out))
    det.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
