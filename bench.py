#!/usr/bin/env python
"""bench.py — Mpixels/s of the feature-detection hot path (Harris + Canny + FHOG) on 4K frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, frames sharded, no
                                                            data-path collective: weak scaling)

A step = one pass of the three detectors over one batch of B synthetic 3840x2160 frames per GPU
(grey plane -> Harris corners and Canny edge map, RGB planes -> FHOG).  The JSON line reports
  value   : whole-job Mpixels/s with the frames already resident in HBM (device-timed, CUDA events)
  e2e     : the same through the public host API (pinned host buffers, H2D + D2H inside the timing)
  roofline: the fused Harris gradient+response kernel, algorithmic 5 B/pixel (u8 in, f32 R out),
            achieved GB/s from CUDA events around back-to-back launches, against MEASURED_PEAKS.json
  cpu_baseline: the reference's own C/C++ (oracle/_ref) on this box's host cores, bounded sample.
`--impl reference` times only the reference CPU implementation (all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY = 3840, 2160
HARRIS_KW = dict(k=0.06, sigma_d=1.0, sigma_i=2.5, threshold=130.0, gaussian=0, gradient=0, measure=0)
CANNY_KW = dict(s=2.0, low_thr=3.0, high_thr=10.0, accGrad=True)
FHOG_KW = dict(cell=8, frp=1, fcp=1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        self.rows, self.times, self.proc, self.index = [], [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            self.times.append(time.perf_counter())

    def count_since(self, t0):
        return sum(1 for t in self.times if t >= t0)

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples taken in [t0, t1] (all samples if None)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for r, t in zip(self.rows, self.times) if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def available_detectors():
    import image_b200
    d = ["harris"]
    try:
        from image_b200 import canny  # noqa: F401
        d.append("canny")
    except ImportError:
        pass
    try:
        from image_b200 import dlib  # noqa: F401
        d.append("fhog")
    except ImportError:
        pass
    return d


# ------------------------------------------------------------------------------------------ reference arm
def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _host_memory_available():
    """Bytes this process may still allocate: MemAvailable, capped by the cgroup limit when there is one."""
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except OSError:
        pass
    try:
        mx = open("/sys/fs/cgroup/memory.max").read().strip()
        if mx != "max":
            cur = int(open("/sys/fs/cgroup/memory.current").read())
            avail = min(avail, int(mx) - cur) if avail is not None else int(mx) - cur
    except (OSError, ValueError):
        pass
    return avail


def run_reference(args, dets):
    """The reference's own CPU implementation (oracle/_ref when built, else the oracle port) on the
    box's host cores, with every core busy: a step is `nf` independent frames (default: up to 64, bounded
    by host memory) whose Harris / Canny / FHOG calls all go through one thread pool.  Canny and FHOG are
    single-threaded in the reference; Harris is OpenMP-parallel inside a frame and gets cores/nf threads."""
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    mem = _host_memory_available()
    nf_mem = max(1, int(mem // (3 << 30))) if mem else 16          # ~1 GB per 4K Canny call, 0.5 GB Harris: keep 3 GB per frame
    # a pooled 4K frame costs about 0.5 s of wall time on a 128-core box: keep the whole run near three minutes
    nf_time = max(4, int(360 / (args.steps + min(args.warmup, 1))))
    nf = max(1, args.ref_frames if args.ref_frames > 0 else min(cores, 64, nf_mem, nf_time))
    omp = max(1, cores // nf)
    os.environ["OMP_NUM_THREADS"] = str(omp)                       # read by libgomp when libref_harris.so is loaded below
    from oracle import pyoracle as po
    from image_b200 import synth
    if po.lib("oracle") is None:
        po.build(ref=False)
        po._cache.clear()
    kind = "reference" if all(po.have_ref(w) for w in ("harris", "canny", "dlib")) else "port"
    impl = "ref" if kind == "reference" else "oracle"
    base = [synth.frame_rgb(2000 + i, NY, NX) for i in range(min(nf, 2))]          # two distinct frames, repeated (as the GPU arm)
    rgb = [base[i % len(base)] for i in range(nf)]
    grey = [(f.astype(np.uint16).sum(axis=2) // 3).astype(np.uint8) for f in base]
    grey = [grey[i % len(base)] for i in range(nf)]

    def step():
        with ThreadPoolExecutor(max_workers=cores) as ex:
            jobs = []
            if "canny" in dets:                                    # longest jobs first
                jobs += [ex.submit(po.canny, g, impl=impl, **CANNY_KW) for g in grey]
            if "harris" in dets:
                jobs += [ex.submit(po.harris_detect, g, precision=0, impl=impl, **HARRIS_KW) for g in grey]
            if "fhog" in dets:
                jobs += [ex.submit(po.fhog, f, impl=impl, **FHOG_KW) for f in rgb]
            for j in jobs:
                j.result()

    for _ in range(args.warmup if args.warmup < 2 else 1):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    mpix = nf * NX * NY / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mpixels/sec (Harris+Canny+HOG) at 4K frames", "value": mpix, "unit": "Mpixels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 (CPU reference)",
        "data": "synthetic",
        "config": {"workload": "+".join(dets) + " @3840x2160", "frames_per_step": nf, "detectors": dets},
        "cpu_baseline": {"value": mpix, "unit": "Mpixels/s", "cores": cores, "cpu_model": _cpu_model(), "omp_num_threads": omp, "kind": kind,
                         "sample": "%d synthetic 4K frame(s) per step (2 distinct), %d steps; all calls of a step in one %d-thread pool, Harris with %d OpenMP thread(s) per frame%s"
                                   % (nf, args.steps, cores, omp, "; Canny FFT through the oracle DFT shim (FFTW3 absent)" if kind == "reference" else "")},
        "e2e": {"value": mpix, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="4K frames per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per reference step (0 = min(cores, 16))")
    ap.add_argument("--detectors", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="device steps only (for ncu launch lists): no per-detector, e2e or CPU legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            dets = args.detectors.split(",") if args.detectors else ["harris", "canny", "fhog"]
            from oracle import pyoracle as po
            if po.lib("oracle") is None:
                po.build(ref=False); po._cache.clear()
            dets = [d for d in dets if (d != "fhog" or hasattr(po.lib("oracle"), "orc_fhog") or po.have_ref("dlib"))]
            run_reference(args, dets)
        return

    import torch
    import torch.distributed as dist
    from image_b200 import _lib, synth
    from image_b200 import harris as H
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    ctx = _lib.context(local)
    dets = args.detectors.split(",") if args.detectors else available_detectors()
    B = args.batch
    W = max(args.warmup, 3)
    K = args.steps

    # ---- one process per GPU: stay on the GPU's NUMA node, so that the pinned frame buffers are local to it
    from image_b200.shard import bind_to_gpu_numa
    aff_prev, aff_new = bind_to_gpu_numa(local) if not os.environ.get("B2F_NO_NUMA_BIND") else (os.sched_getaffinity(0), None)

    # ---- synthetic frames (seeded per rank: frames are independent units, sharded across ranks)
    rgb = synth.batch(synth.frame_rgb, 2000 + 1000 * rank, B, NY, NX, distinct=min(B, 8))      # [B, NY, NX, 3] u8, 8 distinct frames
    grey = (rgb.astype(np.uint16).sum(axis=3) // 3).astype(np.uint8)       # dlib's grey rule (r+g+b)/3, pixel.h:775-783
    h_rgb = torch.from_numpy(rgb).pin_memory()
    h_grey = torch.from_numpy(grey).pin_memory()
    d_rgb = h_rgb.cuda()
    d_grey = h_grey.cuda()
    stream = torch.cuda.Stream()          # a real (non-default) stream: the C ABI launches on it and
    torch.cuda.set_stream(stream)         # the CUDA events below are recorded on the same stream
    sp = stream.cuda_stream

    # ---- device-resident outputs
    d_R = torch.empty((B, NY, NX), dtype=torch.float32, device="cuda")
    cap = 65536
    d_xy = torch.empty((B, cap), dtype=torch.int32, device="cuda")
    d_st = torch.empty((B, cap), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(B, dtype=torch.int32, device="cuda")
    radius = int(2 * HARRIS_KW["sigma_i"] + 0.5)
    if "canny" in dets:
        from image_b200 import canny as Cn
        d_edges = torch.empty((B, NY, NX), dtype=torch.uint8, device="cuda")
        d_nz = torch.empty(B, dtype=torch.int32, device="cuda")
    if "fhog" in dets:
        from image_b200 import dlib as Dl
        hnr, hnc = Dl.fhog_size(NY, NX, **FHOG_KW)
        d_hog = torch.empty((B, hnr, hnc, 31), dtype=torch.float32, device="cuda")

    # The three detectors are independent: each gets its own context (stream + scratch) and they run
    # concurrently, forked from / joined into the timing stream with events.
    ctx_c = _lib.new_context(local) if "canny" in dets else None
    ctx_f = _lib.new_context(local) if "fhog" in dets else None
    st_c = torch.cuda.Stream() if "canny" in dets else None
    st_f = torch.cuda.Stream() if "fhog" in dets else None
    ev_fork = torch.cuda.Event()
    ev_c, ev_f = torch.cuda.Event(), torch.cuda.Event()
    serial = bool(os.environ.get("B2F_BENCH_SERIAL"))

    def step_dev():
        if serial:
            H.harris_corners_dev(d_grey, True, B, NX, NY, cap, d_xy, d_st, d_cnt, d_R=d_R, stream=sp, **HARRIS_KW)
            if "canny" in dets:
                Cn.canny_dev(d_grey, B, NX, NY, d_edges, d_nz, stream=sp, **CANNY_KW)
            if "fhog" in dets:
                Dl.fhog_dev(d_rgb, B, NY, NX, d_hog, stream=sp, **FHOG_KW)
            return
        ev_fork.record(stream)
        if "canny" in dets:
            st_c.wait_event(ev_fork)
            Cn.canny_dev(d_grey, B, NX, NY, d_edges, d_nz, stream=st_c.cuda_stream, ctx=ctx_c, **CANNY_KW)
            ev_c.record(st_c)
        if "fhog" in dets:
            st_f.wait_event(ev_fork)
            Dl.fhog_dev(d_rgb, B, NY, NX, d_hog, stream=st_f.cuda_stream, ctx=ctx_f, **FHOG_KW)
            ev_f.record(st_f)
        # certified path: fused response + error bound -> tolerant NMS -> exact patches -> reference-identical lists
        H.harris_corners_dev(d_grey, True, B, NX, NY, cap, d_xy, d_st, d_cnt, d_R=d_R, stream=sp, **HARRIS_KW)
        if "canny" in dets:
            stream.wait_event(ev_c)
        if "fhog" in dets:
            stream.wait_event(ev_f)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up, then the device-resident timed region (inputs 16 x 8.3 MB grey + 16 x 24.9 MB
    #      RGB per step >> 126 MB L2, so every step streams from HBM)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs a moment before its first line: start it before the warm-up
    for _ in range(W):
        step_dev()
    l0 = lib.b2f_launch_count(ctx)
    l0x = [lib.b2f_launch_count(cx) if cx is not None else 0 for cx in (ctx_c, ctx_f)]
    t_region0 = time.perf_counter()
    ms_total = timed(step_dev, K)
    t_region1 = time.perf_counter()
    launches = int(lib.b2f_launch_count(ctx) - l0)
    for cx, base in zip((ctx_c, ctx_f), l0x):
        if cx is not None:
            launches += int(lib.b2f_launch_count(cx) - base)
    clocks = None
    if rank == 0:
        note = "timed region"
        if sampler.proc and sampler.count_since(t_region0) < 3 and not args.profile_mode:
            # the timed region is shorter than a few 100 ms sampling periods: keep the very same load running
            # (untimed) until enough samples have arrived, and say so
            t_end = time.perf_counter() + 1.5
            while time.perf_counter() < t_end and sampler.count_since(t_region0) < 4:
                step_dev()
                torch.cuda.synchronize()
            t_region1 = time.perf_counter()
            note = "timed region + untimed continuation of the same steps (region shorter than the 100 ms sampling period)"
        clocks = sampler.stop(t_region0, t_region1)
        clocks["window"] = note
    ms_step = ms_total / K
    value = world * B * NX * NY / (ms_step * 1e-3) / 1e6

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"metric": "Mpixels/sec (Harris+Canny+HOG) at 4K frames", "value": value, "unit": "Mpixels/s", "n_gpus": world,
                              "steps": K, "warmup": W, "ms_per_step": ms_step, "profile_mode": True, "gpu_launches": launches, "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-detector device times (explain the headline)
    detail = {}

    def t_of(fn, k=5):
        fn()                                   # first call may grow the context's scratch arena
        torch.cuda.synchronize()
        return timed(fn, k) / k
    th = t_of(lambda: H.harris_response_dev(d_grey, True, B, NX, NY, d_R, stream=sp, **HARRIS_KW))
    tall = t_of(lambda: H.harris_corners_dev(d_grey, True, B, NX, NY, cap, d_xy, d_st, d_cnt, d_R=d_R, stream=sp, **HARRIS_KW))
    tn = max(tall - th, 0.0)
    detail["harris_response_ms"] = th
    detail["harris_certify_nms_ms"] = tn
    detail["harris_cert_stats"] = H.cert_stats(ctx)
    if "canny" in dets:
        detail["canny_ms"] = t_of(lambda: Cn.canny_dev(d_grey, B, NX, NY, d_edges, d_nz, stream=sp, **CANNY_KW))
    if "fhog" in dets:
        detail["fhog_ms"] = t_of(lambda: Dl.fhog_dev(d_rgb, B, NY, NX, d_hog, stream=sp, **FHOG_KW))

    per_det = {"harris": {"device_ms_per_frame": (th + tn) / B, "device_mpix_s": B * NX * NY / ((th + tn) * 1e-3) / 1e6}}
    if "canny" in dets:
        per_det["canny"] = {"device_ms_per_frame": detail["canny_ms"] / B, "device_mpix_s": B * NX * NY / (detail["canny_ms"] * 1e-3) / 1e6}
    if "fhog" in dets:
        per_det["fhog"] = {"device_ms_per_frame": detail["fhog_ms"] / B, "device_mpix_s": B * NX * NY / (detail["fhog_ms"] * 1e-3) / 1e6}
    if "fhog" in dets and rank == 0 and not os.environ.get("B2F_BENCH_NO_SURF"):
        # SURF (SURVEY.md 8d, config C4 recipe: Gaussian blobs) is reported beside the headline, not inside it:
        # host API, pinned frames in, key points + descriptors back on the host.
        try:
            nsf = 4
            blobs = np.stack([synth.frame_blobs(3000 + (i % 2), NY, NX) for i in range(nsf)])
            h_blobs = torch.from_numpy(blobs).pin_memory().numpy()
            Dl.surf_batch(h_blobs)
            t0 = time.perf_counter()
            so = Dl.surf_batch(h_blobs)
            dts = time.perf_counter() - t0
            per_det["surf"] = {"e2e_ms_per_frame": dts / nsf * 1e3, "e2e_mpix_s": nsf * NX * NY / dts / 1e6, "frames": nsf,
                               "points_per_frame": int(np.mean([o["points"] for o in so])), "max_points": 10000, "detection_threshold": 30.0}
        except Exception as ex:
            per_det["surf"] = {"error": str(ex)}

    # ---- roofline of the dominant target kernel: fused Harris gradient+response, 5 B/px algorithmic
    peak, peak_src = peaks()
    alg_bytes = 5.0 * B * NX * NY
    achieved = alg_bytes / (th * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "harris_fused3_kernel<3,7,u8> (one launch: interior and border tiles)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "frac_of_nominal_8000": achieved / 8000.0, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": th,
            "note": "this chain is fp32-FMA / shared-memory bound (>= 110 fp32 lane-ops per 5 algorithmic bytes): "
                    "its HBM fraction cannot exceed ~24 % (DESIGN.md 4.1)"}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            roof["traffic"] = json.load(open(tp)).get("harris_fused2_bytes_per_launch_b%d" % B)
        except Exception:
            pass

    # ---- end to end through the public host API (pinned host in, results back on the host)
    e2e = None
    h2d = d2h = 0
    try:
        np_grey = h_grey.numpy()
        np_rgb = h_rgb.numpy()
        # results land in pinned host buffers too (what a serving loop would keep allocated)
        pin_edges = torch.empty((B, NY, NX), dtype=torch.uint8).pin_memory().numpy() if "canny" in dets else None
        pin_hog = torch.empty((B, hnr, hnc, 31), dtype=torch.float32).pin_memory().numpy() if "fhog" in dets else None

        # The three detector calls are independent API calls; a serving loop issues them from three host
        # threads (ctypes drops the GIL), each on its own context, so uploads, kernels and downloads of the
        # detectors overlap on the full-duplex link.  Inside each call the batch is chunked and pipelined.
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=3)
        serial_e2e = bool(os.environ.get("B2F_BENCH_SERIAL"))

        ctx_h = ctx                 # (contexts are per thread by default: pin the pool's Harris calls to this one)

        def e2e_harris():
            return H.harris_batch_u8(np_grey, cap=cap, raw=True, precision=0, ctx=ctx_h, **HARRIS_KW)

        def e2e_canny():
            return Cn.canny_batch(np_grey, out=pin_edges, ctx=ctx_c, **CANNY_KW)

        def e2e_fhog():
            return Dl.fhog_batch(np_rgb, out=pin_hog, ctx=ctx_f, **FHOG_KW)

        def step_e2e():
            nonlocal h2d, d2h
            jobs = [e2e_harris] + ([e2e_canny] if "canny" in dets else []) + ([e2e_fhog] if "fhog" in dets else [])
            res = [j() for j in jobs] if serial_e2e else [f.result() for f in [pool.submit(j) for j in jobs]]
            hc = res[0][3]
            h2d = np_grey.nbytes
            d2h = int(hc.sum()) * 8 + 4 * B
            if "canny" in dets:
                h2d += np_grey.nbytes
                d2h += pin_edges.nbytes + 4 * B
            if "fhog" in dets:
                h2d += np_rgb.nbytes
                d2h += pin_hog.nbytes
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        ke = max(2, K // 2)
        for _ in range(ke):
            step_e2e()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / ke], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # each detector's host call alone (this rank; the same pinned buffers)
        for name, fn in (("harris", e2e_harris), ("canny", e2e_canny), ("fhog", e2e_fhog)):
            if name in dets and name in per_det:
                fn()
                t1 = time.perf_counter()
                for _ in range(3):
                    fn()
                per_det[name]["e2e_mpix_s"] = B * NX * NY / ((time.perf_counter() - t1) / 3) / 1e6
        e2e = {"value": world * B * NX * NY / float(dt.item()) / 1e6, "unit": "Mpixels/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "harris_batch_u8 / canny_batch / fhog_batch (C ABI *_batch entry points, pinned host buffers; one host thread per detector, batches chunked and pipelined inside each call)"}
    except Exception as ex:   # keep the device-timed line even if the host path fails
        e2e = {"value": None, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)}

    # ---- CPU baseline (rank 0, N=1 only): the reference's own code on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, aff_prev)      # the CPU reference gets every core of the box
            from oracle import pyoracle as po
            if po.lib("oracle") is None:
                po.build(ref=False); po._cache.clear()
            kind = "reference" if all(po.have_ref(w) for w in ("harris", "canny", "dlib")) else "port"
            impl = "ref" if kind == "reference" else "oracle"
            g1, f1 = grey[0], rgb[0]
            t0 = time.perf_counter()
            po.harris_detect(g1, precision=0, impl=impl, **HARRIS_KW)
            t_h = time.perf_counter() - t0
            parts = {"harris_s": t_h}
            tot = t_h
            if "canny" in dets:
                t0 = time.perf_counter(); po.canny(g1, impl=impl, **CANNY_KW); parts["canny_s"] = time.perf_counter() - t0
                tot += parts["canny_s"]
            if "fhog" in dets:
                t0 = time.perf_counter(); po.fhog(f1, impl=impl, **FHOG_KW); parts["fhog_s"] = time.perf_counter() - t0
                tot += parts["fhog_s"]
            cpu = {"value": NX * NY / tot / 1e6, "unit": "Mpixels/s", "cores": os.cpu_count(), "cpu_model": _cpu_model(), "kind": kind,
                   "sample": "1 synthetic 4K frame through %s (Harris OpenMP on all cores; Canny, FHOG single thread as in the reference)" % "+".join(dets),
                   "parts": parts}
            # all cores busy: the reference arm's own step (independent frames in one thread pool), one step, in a
            # child process so that its OpenMP team size and memory stay its own
            try:
                nfr = max(2, min((os.cpu_count() or 2) // 4, 32))
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                    "--ref-frames", str(nfr), "--detectors", ",".join(dets)], capture_output=True, text=True, timeout=600)
                pl = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
                cpu["single_frame_value"] = cpu["value"]
                cpu["value"] = pl["value"]
                cpu["sample"] = pl["cpu_baseline"]["sample"] + " | single frame, one call at a time: %.2f Mpixels/s" % cpu["single_frame_value"]
            except Exception as ex2:
                cpu["pool_error"] = str(ex2)
        except Exception as ex:
            cpu = {"value": None, "unit": "Mpixels/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %s" % ex}

    if rank == 0:
        line = {
            "metric": "Mpixels/sec (Harris+Canny+HOG) at 4K frames", "value": value, "unit": "Mpixels/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (Harris, FHOG) / f64 (Canny)", "data": "synthetic",
            "config": {"workload": "+".join(dets) + " @3840x2160, batch=%d frames per GPU per step" % B,
                       "detectors": dets, "frames_per_gpu": B, "l2": "inputs larger than L2 (%.0f MB per step)" % ((grey.nbytes + rgb.nbytes) / 1e6),
                       "parallelism": "frames sharded over %d GPU(s), no data-path collective" % world,
                       "host_affinity": ("GPU NUMA node, %d cpus" % len(aff_new)) if aff_new else "unchanged (NUMA node of the GPU unknown)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "detail_ms_per_step": detail, "per_detector": per_det,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
