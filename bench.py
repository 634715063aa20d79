#!/usr/bin/env python3
"""bench.py - the headline measurement: output Mpix/s of the reprojection hot path.

    python bench.py --gpus N --steps K --warmup W            (ours: sm_100a kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

Workload (config.workload): BASELINE.json configs[1] - a synthetic 1:6 cubemap with 2048-px
faces reprojected to a full spherical 8192x4096 image with a cubic b-spline (prefiltered),
float RGB. A step is one rendered frame per GPU.

  value     whole-job Mpix/s of the render kernel with the staged source resident in HBM,
            CUDA events around exactly K back-to-back launches, max over ranks
  roofline  algorithmic bytes per launch (output store + the six cube faces) /
            average launch duration of k_render inside that same timed region, against the
            measured copy bandwidth in MEASURED_PEAKS.json
  e2e       the same frames through the host-buffer C ABI a cuda_dispatch::payload() would
            call: eu_source_upload (pinned host raster -> H2D -> IR build -> prefilter) +
            eu_render (kernel -> D2H into a pinned host buffer), wall clock, every step
  cpu_baseline / --impl reference
            the UNMODIFIED reference (oracle/_ref/envutil_ref_fast, built from
            /root/reference by oracle/Makefile: -O3 -march=x86-64-v3, zimt goading back-end,
            zimt thread pool = 2 x hardware threads) running the same job on this box's host
            cores; a step is one process run = its payload() (source build + prefilter +
            render), wall clock minus the time the file shim spent reading/writing rasters.
            Falls back to the C oracle port (OpenMP, all cores) if the binary is absent.
Multi-GPU (weak scaling): rank 0 synthesises the source, broadcasts the raster over NCCL,
every rank stages it and renders its own full frame per step (camera yaw = 360 * rank / N);
no data-path collective inside the timed region.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "output Mpix/s (device-timed)"
UNIT = "Mpix/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="shrink the workload (tests only; invalid as a bench)")
    ap.add_argument("--padded", type=int, default=0, help="1: 16-byte RGB texels in HBM")
    ap.add_argument("--no-tiles", type=int, default=0, help="1: direct-gather kernel (no shared-memory staging)")
    ap.add_argument("--warp-tiles", type=int, default=0,
                    help="1: experimental kernel that stages the gather footprint per warp (k_render_warp)")
    ap.add_argument("--partition", default="frames", choices=["frames", "bands"],
                    help="N > 1: frames = one full frame per rank per step (weak scaling, default); bands = the ranks "
                         "split ONE frame into row bands, gathered on rank 0 over NCCL (strong scaling)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="bands: peer = every rank's render kernel stores its band straight into rank 0's frame over "
                         "NVLink (eu_frame_*; render and gather are one kernel, default); nccl = band buffers + gather")
    ap.add_argument("--narrow-stores", type=int, default=0, help="1: 4-byte pixel stores even into a peer frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (default: min(steps, 5))")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Samples taken inside [t0, t1]; if the region was shorter than one sampling period, the
        samples nearest to it (within 150 ms)."""
        inside = [r for (t, r) in self.rows if t0 <= t <= t1]
        if inside:
            return inside
        near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[:3]
        return [r for (t, r) in near if min(abs(t - t0), abs(t - t1)) < 0.15]

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    @staticmethod
    def summarise(rows):
        if not rows:
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the render kernel, from the
    newest ncu --set full capture summarised under profiles/ (tools/summarise_profile.py)."""
    pd = os.path.join(ROOT, "profiles")
    best = None
    for f in sorted(os.listdir(pd)) if os.path.isdir(pd) else []:
        if f.endswith("_traffic.json"):
            try:
                best = json.load(open(os.path.join(pd, f)))
            except ValueError:
                pass
    return best


def workload(scale):
    from envutil_b200 import workloads
    job, alg = workloads.c2(scale)
    name = ("C2: synthetic 1:6 cubemap %dpx faces -> spherical %dx%d hfov 360, cubic b-spline with prefilter, "
            "float RGB" % (job.facets[0].image.shape[1], job.width, job.height))
    return job, alg, name


# ---------------------------------------------------------------------------------------------
# CPU reference leg
def run_reference_cpu(job, steps, warmup, workdir):
    """Times the unmodified reference on this box's host cores. Returns a dict."""
    from envutil_b200 import euf
    exe = os.path.join(ROOT, "oracle", "_ref", "envutil_ref_fast")
    isa = "-march=x86-64-v3"
    try:  # the AVX-512 build of the same sources where the host has it (a zimt vector = one zmm register)
        flags = open("/proc/cpuinfo").read()
        exe512 = exe + "512"
        if all((" " + f) in flags for f in ("avx512f", "avx512vl", "avx512bw", "avx512dq", "avx512cd")) and os.path.exists(exe512):
            exe, isa = exe512, "-march=x86-64-v4"
    except OSError:
        pass
    mpix = job.width * (job.height or job.width) / 1e6
    ncores = os.cpu_count() or 1
    if os.path.exists(exe):
        paths = []
        for i, f in enumerate(job.facets):
            p = os.path.join(workdir, "facet%d.euf" % i)
            euf.write_euf(p, f.image)
            paths.append(p)
        cmd = [exe, "-v"] + job.cli_args(paths, os.path.join(workdir, "out.euf"))
        env = dict(os.environ, EUSHIM_NOWRITE="1")
        payload_s, render_s = [], []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            wall = time.perf_counter() - t0
            if r.returncode != 0 and i == 0 and exe.endswith("512"):
                # the AVX-512 build does not run on this host after all: the AVX2 build of the same sources
                exe, isa = exe[:-3], "-march=x86-64-v3"
                cmd[0] = exe
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True, env=env)
                wall = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError("reference failed: " + r.stderr[-2000:])
            rd = re.findall(r"eushim: read time ([0-9.eE+-]+) ms", r.stdout)
            wr = re.findall(r"eushim: write time ([0-9.eE+-]+) ms", r.stdout)
            fr = re.findall(r"frame rendering time: ([0-9]+) ms", r.stdout)
            io_s = (float(rd[-1]) if rd else 0.0) / 1e3 + (float(wr[-1]) if wr else 0.0) / 1e3
            if i >= warmup:
                payload_s.append(wall - io_s)
                if fr:
                    render_s.append(max(float(fr[-1]) / 1e3 - (float(wr[-1]) if wr else 0.0) / 1e3, 1e-6))
        t = float(np.mean(payload_s))
        return {"value": mpix / t, "unit": UNIT, "cores": ncores, "threads": 2 * ncores, "kind": "reference",
                "sample": "%d full frames (%.1f Mpix each), one process run per frame = payload(): source build + "
                          "prefilter + render, wall clock minus raster file I/O" % (steps, mpix),
                "ms_per_step": t * 1e3,
                "render_only_mpix_s": (mpix / float(np.mean(render_s))) if render_s else None,
                "build": "oracle/_ref/%s: unmodified reference sources, g++ -O3 %s, zimt goading back-end"
                         % (os.path.basename(exe), isa)}
    # fallback: the C oracle port, a band of rows sized for a few seconds
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    st = job.structs()
    t0 = time.perf_counter()
    hs = harness.oracle_sources(job, st)
    stage_s = time.perf_counter() - t0
    rows = max(8, min(st[0].height, int(2e6 / st[0].width)))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        harness.oracle_render(job, rows=(0, rows), sources=hs)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    t_render_full = float(np.mean(ts)) * st[0].height / rows
    t = stage_s + t_render_full
    return {"value": mpix / t, "unit": UNIT, "cores": ncores, "threads": ncores, "kind": "port",
            "sample": "oracle port (OpenMP): source staging once + %d of %d rows per step, scaled to the frame"
                      % (rows, st[0].height),
            "ms_per_step": t * 1e3, "render_only_mpix_s": mpix / t_render_full, "build": "oracle/liboracle.so"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    job, alg, name = workload(args.scale)
    with tempfile.TemporaryDirectory(prefix="eubench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        steps = max(1, min(args.steps, 5))
        warm = max(1, min(args.warmup, 1))
        cb = run_reference_cpu(job, steps, warm, d)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "frames_per_step": 1, "note": "CPU reference runs once on rank 0's host "
                       "cores; steps/warmup bounded to %d/%d process runs" % (steps, warm)},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "threads", "kind", "sample", "build",
                                                "render_only_mpix_s")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from envutil_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: rank 0 synthesises, NCCL broadcasts the raster (the only exchange step) ----
    t_setup = time.perf_counter()
    if rank == 0:
        job, alg, name = workload(args.scale)
        src = torch.from_numpy(job.facets[0].image)
        meta = torch.tensor(list(src.shape), dtype=torch.int64, device=dev)
    else:
        meta = torch.zeros(3, dtype=torch.int64, device=dev)
    bcast_ms = 0.0
    if world > 1:
        dist.broadcast(meta, 0)
    shape = tuple(int(v) for v in meta.tolist())
    if rank == 0:
        d_src = src.to(dev)
    else:
        d_src = torch.empty(shape, dtype=torch.float32, device=dev)
    if world > 1:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(d_src, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            from envutil_b200.job import FacetSpec, Job
            from envutil_b200 import workloads
            face = shape[1]
            job = Job([FacetSpec(None, "cubemap", 90.0, width=face, height=6 * face, nchannels=3)], "spherical", 360.0,
                      4 * face, 2 * face, degree=3, name="C2")
            alg = job.width * job.height * workloads.RGB + 6 * face * face * workloads.RGB
            name = ""
    h_src = torch.empty(shape, dtype=torch.float32).pin_memory()
    h_src.copy_(d_src)
    torch.cuda.synchronize()
    bands_mode = args.partition == "bands" and world > 1
    if not bands_mode:
        job.yaw = 360.0 * rank / world  # every rank renders its own view of the same environment
    job.padded, job.no_tiles = bool(args.padded), bool(args.no_tiles)
    job.warp_tiles = bool(args.warp_tiles)
    job.narrow_stores = bool(args.narrow_stores)
    eng = Engine(local)
    st = job.structs(eng.lib)
    t, fa, o, taps, ntaps = st
    H, W, C = t.height, t.width, t.nchannels
    mpix = W * H / 1e6
    stream = torch.cuda.current_stream().cuda_stream
    hs = eng.stage_device(job, [d_src.data_ptr()], st, stream=stream, padded=bool(args.padded))
    eng.release(hs)  # the first staging grows the device pool; report the steady state
    hs = eng.stage_device(job, [d_src.data_ptr()], st, stream=stream, padded=bool(args.padded))
    stage_ms = eng.last_stage_timing[0].render_ms
    import ctypes as _C
    ir_bytes = 4 * int(eng.lib.eu_source_container_floats(hs[0], (_C.c_int32 * 4)()))  # the staged cubemap IR
    stage_launches = eng.last_stage_timing[0].launches
    from envutil_b200 import bands as eu_bands
    row0, row1 = eu_bands.band(H, world, rank) if bands_mode else (0, H)
    peer = None
    if bands_mode and args.gather == "peer":
        peer = eu_bands.PeerFrame(eng.lib, dist, H, W, C, rank, world)
        out_ptr = peer.band_ptr(row0)
        d_out = None
    else:
        d_out = torch.empty((row1 - row0, W, C), dtype=torch.float32, device=dev)
        out_ptr = d_out.data_ptr()
    setup_s = time.perf_counter() - t_setup

    # ---- device-timed render: W warm-up launches, then exactly K, events on the launch stream
    for _ in range(max(args.warmup, 3)):
        eng.render_rows(job, hs, st, row0, row1, out_ptr, stream, timed=False)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:  # nvidia-smi takes a moment to deliver its first sample
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.02)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = eng.launches
    tw0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        eng.render_rows(job, hs, st, row0, row1, out_ptr, stream, timed=False)
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launches - launches_before
    tm = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_per_step = float(tm.item()) / args.steps
    clocks = None
    if sampler:
        time.sleep(0.15)
        clocks = ClockSampler.summarise(sampler.window(tw0, tw1))
        # the sampler covers the device-timed region only: polling NVML every 20 ms contends with the
        # driver and disturbs the asynchronous e2e leg below (measured: 9.4 ms/frame without it, 12-49 ms with)
        sampler.stop()
        sampler = None
    gather_ms = 0.0
    if peer is not None:  # the timed launches already left every band in rank 0's frame (barrier above)
        checksum = float(peer.as_tensor()[::64, ::64].double().sum().item()) if rank == 0 else 0.0
    else:
        checksum = float(d_out[::64, ::64].double().sum().item())
    if bands_mode and peer is None:  # output bands gathered on rank 0 (NCCL), timed on its own: not part of `value`
        full = eu_bands.gather_bands(d_out, H, world, rank, dist)  # first use sets up the NCCL channels
        del full
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = eu_bands.gather_bands(d_out, H, world, rank, dist)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        if rank == 0:
            assert tuple(full.shape) == (H, W, C)
            checksum = float(full[::64, ::64].double().sum().item())
        del full

    # ---- e2e: host buffers through the C ABI, H2D + staging + render + D2H every step ----
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
    peer_ok = None
    if bands_mode:  # the e2e leg renders full frames through eu_render; compare against a full device frame
        d_out = torch.empty((H, W, C), dtype=torch.float32, device=dev)
        eng.render_rows(job, hs, st, 0, H, d_out.data_ptr(), stream, timed=False)
        torch.cuda.synchronize()
        if peer is not None:
            if rank == 0:  # the frame assembled by the ranks' peer stores equals one GPU's full render
                peer_ok = bool(torch.equal(peer.as_tensor(), d_out))
            barrier()
            if rank != 0:
                peer.close()
            barrier()
            if rank == 0:
                peer.close()
    h_out = torch.empty((H, W, C), dtype=torch.float32).pin_memory()
    import ctypes as Ct
    from envutil_b200 import capi

    def e2e_step():
        tmu, tmr = capi.Timing(), capi.Timing()
        h = capi.SourceH()
        capi.check(eng.lib.eu_source_upload(None, Ct.byref(fa[0]), Ct.byref(o), h_src.data_ptr(), Ct.byref(h),
                                            Ct.byref(tmu)), eng.lib)
        one = (capi.SourceH * 1)(h)
        capi.check(eng.lib.eu_render(Ct.byref(t), Ct.byref(o), 1, fa, one, taps, ntaps, h_out.data_ptr(),
                                     Ct.byref(tmr)), eng.lib)
        capi.check(eng.lib.eu_source_release(h), eng.lib)
        return tmu, tmr

    # the host side of this leg is a Python loop over ctypes calls: keep the cyclic garbage collector out
    # of the timed regions (a generation-2 pass over torch's heap costs tens of ms and used to land on
    # the same frame of every run)
    import gc
    gc.collect()
    gc.disable()
    e2e_step()
    # both legs are repeated three times and the fastest repetition is reported: the PCIe links and the
    # host memory of the box are shared with whatever runs on its other GPUs, and a single burst of that
    # traffic otherwise decides the number (observed: isolated 25-130 ms gaps between results)
    REPS = 3
    e2e_blocking_s, parts = None, None
    for _ in range(REPS):
        barrier()
        t0 = time.perf_counter()
        these = []
        for _ in range(e2e_steps):
            these.append(e2e_step())
        barrier()
        dt = time.perf_counter() - t0
        if e2e_blocking_s is None or dt < e2e_blocking_s:
            e2e_blocking_s, parts = dt, these
    # pipelined: the same per-step work (H2D of the raster, staging, render, D2H of the frame) through
    # eu_source_upload_async / eu_render_async / eu_job_wait with up to three jobs in flight, so the
    # upload of step n+1 overlaps the download of step n
    DEPTH = 3
    ring = [h_out] + [torch.empty((H, W, C), dtype=torch.float32).pin_memory() for _ in range(DEPTH - 1)]
    pipe_steps = max(e2e_steps, 8 * DEPTH)

    finish_stamps = []

    def pipelined(n_steps):
        pending = []
        del finish_stamps[:]
        for i in range(n_steps):
            pending.append(eng.submit(job, st, [h_src.data_ptr()], ring[i % DEPTH].data_ptr()))
            if len(pending) >= DEPTH:
                eng.finish(pending.pop(0))
                finish_stamps.append(time.perf_counter())
        while pending:
            eng.finish(pending.pop(0))
            finish_stamps.append(time.perf_counter())

    # warm-up until the stream-ordered pool has enough containers cycling between the upload and the
    # staging stream (the first passes still grow it: 321 MB from the driver per job, tens of ms each)
    pipelined(4 * DEPTH)
    e2e_s, best_stamps, best_t0 = None, None, None
    for _ in range(REPS):
        barrier()
        t0 = time.perf_counter()
        pipelined(pipe_steps)
        barrier()
        dt = time.perf_counter() - t0
        if e2e_s is None or dt < e2e_s:
            e2e_s, best_stamps, best_t0 = dt, list(finish_stamps), t0
    finish_stamps[:] = best_stamps
    t0 = best_t0
    gc.enable()
    te = torch.tensor([e2e_s / pipe_steps, e2e_blocking_s / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0].item()) * 1e3
    e2e_blocking_ms = float(te[1].item()) * 1e3
    e2e_steps = pipe_steps
    h_out = ring[(pipe_steps - 1) % DEPTH]
    e2e_ok = bool(np.array_equal(h_out[::97, ::89].numpy(), d_out[::97, ::89].cpu().numpy()))
    if sampler:
        sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peak()
        tr = measured_traffic() if (args.scale == 1 and not args.padded and not args.no_tiles and not args.warp_tiles
                                    and not bands_mode) else None
        alg_launch = alg // world if bands_mode else alg  # a band touches its share of output and source
        achieved = alg_launch / (ms_per_step * 1e-3) / 1e9  # per launch = per GPU
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            with tempfile.TemporaryDirectory(prefix="eubench_",
                                             dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
                try:
                    cb = run_reference_cpu(job, 2, 1, d)
                    cpu = {k: cb[k] for k in ("value", "unit", "cores", "threads", "kind", "sample", "build",
                                              "render_only_mpix_s")}
                except Exception as e:  # the baseline is context; never let it sink the bench line
                    cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                           "sample": "failed: %s" % str(e)[:200]}
        line = {
            "metric": METRIC, "value": (1 if bands_mode else world) * mpix / (ms_per_step * 1e-3), "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if bands_mode else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": name, "frames_per_step": 1 if bands_mode else world,
                       "partition": "row bands of one frame, gathered on rank 0" if bands_mode else "one frame per rank", "out_mpix_per_frame": mpix,
                       "l2": "inputs_larger_than_l2 (321 MB source IR + 403 MB output per frame vs 126 MB L2)",
                       "texel_layout": "float4-padded" if args.padded else "interleaved-rgb",
                       "gather": "direct (L1)" if args.no_tiles else
                                 "footprint staged in shared memory per warp by cp.async.bulk (experimental)" if args.warp_tiles else
                                 "footprint staged in shared memory by cp.async.bulk",
                       "arithmetic": capi.ARITHMETIC,  # "contracted" only with EU_ARITHMETIC=contracted (opt-in build)
                       "parity": "bit-exact vs pinned-math reference build (tests/)" if capi.ARITHMETIC == "exact" else
                                 "window evaluation with fused multiply-adds: indices identical, values within 3.5e-7 "
                                 "relative (RMS 5.8e-8) of the pinned-math reference build on this workload, 2.6e-6 "
                                 "(hdr_merge 1.6e-5) over the 98 small jobs (tests/test_contracted.py, DESIGN.md 2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": tr["traffic_bytes"] if tr else None,
                         "traffic_source": tr["source"] if tr else None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_launch,
                         "kernel": "k_render<3,...>" if args.no_tiles else "k_render_warp<3,...>" if args.warp_tiles else "k_render_tiled<3,...>", "frac_of_8TBs_spec": achieved / 8000.0},
            "e2e": {"value": world * mpix / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h_src.numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4),
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "scope": "every step: H2D of the 302 MB raster + cubemap IR + prefilter + render + D2H of the 403 MB "
                             "frame, pinned host buffers, wall clock; eu_source_upload_async / eu_render_async / "
                             "eu_job_wait with 3 jobs in flight (upload of step n+1 overlaps download of step n); "
                             "fastest of 3 repetitions of the whole %d-frame run" % pipe_steps,
                    "blocking": {"value": world * mpix / (e2e_blocking_ms * 1e-3), "ms_per_step": e2e_blocking_ms,
                                 "scope": "eu_source_upload + eu_render, one blocking call pair per step"},
                    "ms_between_results": [round((b - a) * 1e3, 2) for a, b in zip([t0] + finish_stamps, finish_stamps)],
                    "median_ms_between_results": float(np.median(np.diff(np.array([t0] + finish_stamps)))) * 1e3,
                    "matches_device_path": e2e_ok,
                    "breakdown_ms": {"h2d": float(np.mean([p[0].h2d_ms for p in parts])),
                                     "staging_kernels": float(np.mean([p[0].render_ms for p in parts])),
                                     "render_kernel": float(np.mean([p[1].render_ms for p in parts])),
                                     "d2h": float(np.mean([p[1].d2h_ms for p in parts]))}},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "staging": {"ms": stage_ms, "launches": stage_launches,
                        "what": "cubemap IR build + support fill + per-section prefilter (device-timed, outside value)",
                        # the IR build reads the raster and writes the IR, the prefilter reads
                        # and writes the IR once per axis
                        "algorithmic_bytes": int(h_src.numel() * 4 + ir_bytes + 4 * ir_bytes),
                        "achieved_gbs": (h_src.numel() * 4 + 5 * ir_bytes) / (stage_ms * 1e-3) / 1e9,
                        "frac_measured_peak": (h_src.numel() * 4 + 5 * ir_bytes) / (stage_ms * 1e-3) / 1e9 / peak},
            "multi_gpu": {"broadcast_ms": bcast_ms, "gather_ms": gather_ms, "collectives_in_timed_region": 0,
                          "gather": (("peer stores over NVLink into rank 0's frame inside the timed render kernels "
                                      "(eu_frame_*), no band buffers") if peer_ok is not None else
                                     ("NCCL gather of band buffers, timed separately" if bands_mode else None)),
                          "peer_frame_equals_single_gpu_render": peer_ok},
            "cpu_baseline": cpu, "checksum": checksum, "setup_s": setup_s,
        }
        print(json.dumps(line))
    eng.release(hs)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
