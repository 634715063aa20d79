#!/usr/bin/env python3
"""bench.py - the headline measurement: output Mpix/s of the reprojection hot path.

    python bench.py --gpus N --steps K --warmup W            (ours: sm_100a kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

N = 1  workload (config.workload): BASELINE.json configs[1] - a synthetic 1:6 cubemap with 2048-px faces
       reprojected to a full spherical 8192x4096 image with a cubic b-spline (prefiltered), float RGB; a step is
       one rendered frame. The line also carries `configs`: every other BASELINE config measured the same way in
       the same run (device-timed render, roofline, staging), and the configs[4] pipeline on this one GPU.
N > 1  workload: BASELINE.json configs[4] ("... at 1/2/4/8 GPUs") - the PTO stitch of 6 positions x 3 exposure
       brackets into a 16384x8192 panorama, ONE frame split into row bands over the ranks (strong scaling;
       envutil_b200/c5.py). A step is the whole two-stage pipeline. `--workload c2 --partition frames|bands`
       keeps round 1's replica frames / C2 bands.

  value     whole-job Mpix/s, CUDA events around exactly K steps with all inputs resident in HBM, max over ranks
  roofline  algorithmic bytes per launch of the dominant kernel / its average duration inside that timed region,
            against the measured copy bandwidth in MEASURED_PEAKS.json; `secondary` = the instruction-issue bound
            of the same kernel from its ncu instruction count (profiles/)
  e2e       the same work through the host-buffer C ABI: page-locked host rasters -> H2D -> staging -> render ->
            D2H into a page-locked host frame, wall clock, every step (N > 1: every rank moves only its own part
            of the sources and its own band, over its own PCIe link, into one frame in shared host memory)
  cpu_baseline / --impl reference
            the UNMODIFIED reference (oracle/_ref/envutil_ref_fast[512], built from /root/reference by
            oracle/Makefile: -O3, zimt goading back-end - highway is not in this image -, zimt thread pool =
            2 x hardware threads) running the same job on this box's host cores; a step is one process run = its
            payload() (source build + prefilter + render), wall clock minus the time the file shim spent on rasters.
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
DEFAULT_ARITHMETIC = "exact"
UNIT = "Mpix/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="shrink the workload (tests only; invalid as a bench)")
    ap.add_argument("--padded", type=int, default=-1, help="RGB texels in HBM: -1 = the library's rule (16-byte for bilinear jobs, 12-byte for the cubic headline), 1 = 16-byte, 0 = 12-byte")
    ap.add_argument("--no-tiles", type=int, default=0, help="1: direct-gather kernel (no shared-memory staging)")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c5"],
                    help="auto: C2 on one GPU (the headline config), the C5 row-band pipeline on N > 1")
    ap.add_argument("--arithmetic", default=DEFAULT_ARITHMETIC, choices=["exact", "contracted"],
                    help="render kernels: exact = no contraction (bit-identical to the reference's parity build); contracted = "
                         "fused multiply-adds in the window evaluation (EU_OPT_CONTRACTED; parity in profiles/)")
    ap.add_argument("--configs", default="C1,C3a,C3b,C4,C5",
                    help="N = 1: the other BASELINE configs measured beside the headline ('' = none)")
    ap.add_argument("--budget-s", type=float, default=240.0,
                    help="N = 1: configs that would start after this many seconds of run time are skipped (and listed)")
    ap.add_argument("--c5-plan", default="needed", choices=["needed", "full"],
                    help="C5 stage A: merge only the texels stage B samples (default) or every texel of every position")
    ap.add_argument("--partition", default="frames", choices=["frames", "bands"],
                    help="--workload c2, N > 1: frames = one full frame per rank per step (replicas, weak scaling); bands "
                         "= the ranks split ONE C2 frame into row bands (strong scaling)")
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
# evidence files under profiles/ (committed; produced on a B200 by tools/full_parity.py and tools/gpu_profile.sh)
def _profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except ValueError:
            pass
    return None


def parity_record(config, arithmetic):
    """max / RMS of the CUDA path against both builds of the reference for one config at FULL size, whole frame
    (tools/full_parity.py on the B200 box; bench.py itself never runs the checker outside its cpu_baseline leg)."""
    d = _profile_json("r02_parity_full_%s.json" % arithmetic)
    if not d:
        return {"source": None, "note": "profiles/r02_parity_full_%s.json not present" % arithmetic}
    for rec in d.get("configs", []):
        if rec.get("config") == config:
            out = {"source": "profiles/r02_parity_full_%s.json (tools/full_parity.py, whole frame at full size)" % arithmetic,
                   "eps": d.get("eps"), "tolerance": d.get("tolerance")}
            vp = rec.get("vs_pinned", {})
            out["vs_pinned"] = {"max": vp.get("max_rel"), "rms": vp.get("rms_rel"), "n_diff": vp.get("n_diff"), "n": vp.get("n")}
            vl = rec.get("vs_libm")
            if vl:
                m = vl.get("masked", {})
                out["vs_libm"] = {"max": vl.get("max_rel"), "rms": vl.get("rms_rel"), "n_beyond_1e-5": vl.get("n_beyond_1e-5"),
                                  "masked": m.get("pixels"), "max_outside_tie_band": m.get("max_rel")}
            rs = rec.get("ref_self")
            if rs:
                out["reference_builds_apart"] = {"max": rs.get("max_rel"), "rms": rs.get("rms_rel")}
            if "round_trip" in rec:
                out["round_trip"] = rec["round_trip"]
            return out
    return {"source": None, "note": "no record for %s" % config}


def kernel_record(config, arithmetic):
    """ncu figures of the config's render kernel (one launch, --set full): DRAM bytes, instructions, L2 hit rate."""
    d = _profile_json("r02_kernels.json") or {}
    return d.get("%s/%s" % (config, arithmetic)) or d.get(config)


def roofline_of(alg_bytes, ms, peak, peak_src, config, arithmetic, sm_mhz=1965.0, launches=1):
    """alg_bytes and ms cover `launches` equal launches of the config's kernel (C5-A: one per position); the ncu
    record holds ONE launch. launches = 0: the timed launches are not the recorded one (no ncu figures)."""
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    k = kernel_record(config, arithmetic) if launches else None
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": k.get("dram_bytes") * launches if k else None, "traffic_source": k.get("source") if k else None,
         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes), "frac_of_8TBs_spec": achieved / 8000.0}
    if launches > 1:
        r["launches"] = launches
    if k and k.get("inst_executed"):
        # the kernel's instruction stream at one warp instruction per scheduler per cycle: 148 SMs x 4 schedulers
        floor_ms = launches * k["inst_executed"] / (148 * 4 * sm_mhz * 1e6) * 1e3
        r["secondary"] = {"bound": "issue", "frac": floor_ms / ms, "floor_ms": floor_ms, "inst_executed": k["inst_executed"],
                          "l2_hit_pct": k.get("l2_hit_pct"), "issue_active_pct": k.get("issue_active_pct"),
                          "kernel": k.get("kernel"), "source": k.get("source")}
    return r


def numa_setup(local):
    """Pin this rank to the CPUs next to its GPU (NVML affinity) before any page-locked buffer is allocated, so that
    first-touch places the buffers on the GPU's NUMA node. Returns what was done, for the bench line."""
    info = {"cpus_before": len(os.sched_getaffinity(0))}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        info.update({"gpu_cpus": len(cpus), "cpus_after": len(os.sched_getaffinity(0))})
        try:
            info["numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            pass
    except Exception as e:  # affinity is an optimisation, never a requirement
        info["error"] = str(e)[:120]
    return info


def time_launches(torch, fn, steps, warmup=3, flush=None):
    """Average device time of fn() in ms: CUDA events on the launching stream around `steps` back-to-back calls
    (flush: an L2-sized buffer cleared before every call, each call timed on its own)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush is None:
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def bench_other_configs(args, eng, torch, peak, peak_src, t_run0):
    """Every BASELINE config besides the headline, in this run: device-timed render (inputs resident in HBM, larger
    than L2 or L2 flushed), roofline against exact algorithmic bytes, staging time, parity record of the config."""
    from envutil_b200 import c5, synth, workloads
    from envutil_b200.job import FacetSpec, Job
    want = [c for c in args.configs.split(",") if c]
    out, skipped = [], []
    contracted = args.arithmetic == "contracted"
    stream = torch.cuda.current_stream().cuda_stream
    steps = max(3, min(args.steps, 10))

    def over_budget(name):
        if time.perf_counter() - t_run0 > args.budget_s:
            skipped.append(name)
            return True
        return False

    def single(job, alg, name, flush=None, keep=None, dev_src=None):
        job.contracted = contracted
        st = job.structs(eng.lib)
        t = st[0]
        if dev_src is not None:
            hs = eng.stage_device(job, [dev_src.data_ptr()], st, stream=stream)
        else:
            hs = eng.stage(job, st)
        stage_ms = sum(tm.render_ms for tm in eng.last_stage_timing)
        d_out = torch.empty((t.height, t.width, t.nchannels), dtype=torch.float32, device="cuda")
        ms = time_launches(torch, lambda: eng.render_rows(job, hs, st, 0, t.height, d_out.data_ptr(), stream, timed=False),
                           steps, 3, flush)
        eng.release(hs)
        mpix = t.width * t.height / 1e6
        rec = {"config": name, "workload": job.name, "out": "%dx%d" % (t.width, t.height), "value": mpix / (ms * 1e-3),
               "unit": UNIT, "ms": ms, "steps": steps,
               "l2": "flushed between launches" if flush is not None else "inputs larger than L2",
               "roofline": roofline_of(alg, ms, peak, peak_src, name, args.arithmetic),
               "staging": {"ms": stage_ms}, "parity": parity_record(name, args.arithmetic)}
        out.append(rec)
        return d_out if keep else None

    if "C1" in want and not over_budget("C1"):
        job, alg = workloads.c1(args.scale)
        job.name = "C1: lat/lon 4096x2048 -> rectilinear 1920x1080 hfov 90, bilinear"
        single(job, alg, "C1", flush=torch.empty(256 << 20, dtype=torch.uint8, device="cuda"))
    if ("C3a" in want or "C3b" in want) and not over_budget("C3"):
        job, alg = workloads.c3a(args.scale)
        job.name = "C3a: lat/lon 16384x8192 -> biatan6 cubemap 4096px faces, bilinear"
        cube = single(job, alg, "C3a", keep=True)
        del job
        if "C3b" in want:
            face = cube.shape[1]
            job2 = Job([FacetSpec(None, "biatan6", 90.0, width=face, height=6 * face, nchannels=3)], "spherical", 360.0,
                       4 * face, 2 * face, degree=1, name="C3b: that biatan6 cubemap -> lat/lon 16384x8192, bilinear")
            alg2 = job2.width * job2.height * workloads.RGB + 6 * face * face * workloads.RGB
            if face == 4096:
                alg2 = job2.width * job2.height * workloads.RGB + workloads.EXACT_TOUCHED["C3b"] * workloads.RGB
            single(job2, alg2, "C3b", dev_src=cube)
        del cube
    if "C4" in want and not over_budget("C4"):
        job, alg = workloads.c4(args.scale)
        job.name = "C4: lat/lon 8192x4096 -> fisheye 4096x4096 hfov 180, twine 4 (16 sub-rays per pixel), bilinear"
        single(job, alg, "C4")
    if "C5" in want and not over_budget("C5"):
        (w, h), (W, H) = c5.sizes(args.scale)
        checks = {}
        for plan in ("full", "needed"):
            pl = c5.Pipeline(eng, torch, 0, 1, args.scale, plan=plan, contracted=contracted)
            pl.upload()
            a_ms = time_launches(torch, pl.stage_a, 3, 2)
            pl.stage_b_staging()
            sb_ms = time_launches(torch, pl.stage_b_staging, 3, 1)
            b_ms = time_launches(torch, pl.stage_b, 3, 2)
            pipe_ms = time_launches(torch, pl.step_device, 3, 1)
            torch.cuda.synchronize()
            checks[plan] = float(pl.d_band[0][::61, ::67].double().sum().item())
            mp_a = c5.stage_a_pixels(pl.rects) * c5.POSITIONS / 1e6
            if plan == "full":
                out.append({"config": "C5A", "workload": "C5 stage A: 6 x hdr_merge (--single 0) of 3 brackets %dx%d, every texel "
                            "merged (as the reference runs it)" % (w, h), "out": "6 x %dx%d" % (w, h),
                            "value": mp_a / (a_ms * 1e-3), "unit": UNIT, "ms": a_ms, "steps": 3, "l2": "inputs larger than L2",
                            "roofline": roofline_of(pl.stage_a_alg_bytes(), a_ms, peak, peak_src, "C5A", args.arithmetic,
                                                    launches=c5.POSITIONS),
                            "staging": {"ms": None, "note": "bilinear sources: brace only, inside the upload"},
                            "parity": parity_record("C5A", args.arithmetic)})
                out.append({"config": "C5B", "workload": "C5 stage B: voronoi panorama of the 6 merged rasters -> spherical %dx%d"
                            % (W, H), "out": "%dx%d" % (W, H), "value": W * H / 1e6 / (b_ms * 1e-3), "unit": UNIT, "ms": b_ms,
                            "steps": 3, "l2": "inputs larger than L2",
                            "roofline": roofline_of(pl.alg_b_full, b_ms, peak, peak_src, "C5B", args.arithmetic),
                            "staging": {"ms": sb_ms, "what": "brace of the six merged rasters"},
                            "parity": parity_record("C5B", args.arithmetic)})
            rec = {"config": "C5 pipeline (%s)" % plan, "out": "%dx%d" % (W, H), "value": W * H / 1e6 / (pipe_ms * 1e-3),
                   "unit": UNIT, "ms": pipe_ms, "stage_a_ms": a_ms, "stage_b_staging_ms": sb_ms, "stage_b_ms": b_ms,
                   "stage_a_mpix": mp_a, "h2d_bytes": pl.h2d_bytes, "checksum": checks[plan],
                   "workload": "configs[4] on ONE GPU: stage A + brace + stage B, brackets resident in HBM; plan '%s' = %s" % (
                       plan, "stage A merges every texel of every position" if plan == "full" else
                       "stage A merges only the rectangles of each position that stage B samples (49 % of the columns)")}
            out.append(rec)
            pl.close()
            del pl
        out[-1]["checksum_equals_full_merge"] = checks["needed"] == checks["full"]
    return out, skipped


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
    numa = numa_setup(local)
    from envutil_b200 import synth
    synth.WORKERS = max(1, min(16, len(os.sched_getaffinity(0)) // max(1, world)))
    t_run0 = time.perf_counter()
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
    job.padded, job.no_tiles = (None if args.padded < 0 else bool(args.padded)), bool(args.no_tiles)
    job.contracted = args.arithmetic == "contracted"
    job.narrow_stores = bool(args.narrow_stores)
    eng = Engine(local)
    st = job.structs(eng.lib)
    t, fa, o, taps, ntaps = st
    H, W, C = t.height, t.width, t.nchannels
    mpix = W * H / 1e6
    stream = torch.cuda.current_stream().cuda_stream
    hs = eng.stage_device(job, [d_src.data_ptr()], st, stream=stream)
    eng.release(hs)  # the first staging grows the device pool; report the steady state
    hs = eng.stage_device(job, [d_src.data_ptr()], st, stream=stream)
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
    rep_times = []
    for _ in range(REPS):
        barrier()
        t0 = time.perf_counter()
        pipelined(pipe_steps)
        barrier()
        dt = time.perf_counter() - t0
        rep_times.append(dt)
        if e2e_s is None or dt < e2e_s:
            e2e_s, best_stamps, best_t0 = dt, list(finish_stamps), t0
    finish_stamps[:] = best_stamps
    t0 = best_t0
    gc.enable()
    rep_times.sort()
    e2e_median_s = rep_times[len(rep_times) // 2]
    te = torch.tensor([e2e_s / pipe_steps, e2e_blocking_s / e2e_steps, e2e_median_s / pipe_steps], dtype=torch.float64,
                      device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0].item()) * 1e3
    e2e_blocking_ms = float(te[1].item()) * 1e3
    e2e_median_ms = float(te[2].item()) * 1e3
    e2e_steps = pipe_steps
    h_out = ring[(pipe_steps - 1) % DEPTH]
    e2e_ok = bool(np.array_equal(h_out[::97, ::89].numpy(), d_out[::97, ::89].cpu().numpy()))
    if sampler:
        sampler.stop()

    other, skipped = [], []
    if rank == 0 and world == 1 and args.configs:
        eng.release(hs)
        hs = None
        del d_out
        torch.cuda.empty_cache()
        pk, pk_src = measured_peak()
        other, skipped = bench_other_configs(args, eng, torch, pk, pk_src, t_run0)
    if rank == 0:
        peak, peak_src = measured_peak()
        tr = measured_traffic() if (args.scale == 1 and args.padded != 1 and not args.no_tiles
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
        roof = roofline_of(alg_launch, ms_per_step, peak, peak_src, "C2", args.arithmetic,
                           clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0)
        if roof["traffic"] is None and tr:
            roof["traffic"], roof["traffic_source"] = tr["traffic_bytes"], tr["source"]
        roof["kernel"] = "k_render<3,...>" if args.no_tiles else "k_render_tiled<3,...>"
        if cpu and cpu.get("build"):
            cpu["build"] += ("; highway / Vc are not in this image, so this is zimt's plain-loop ('goading') back-end - the "
                             "reference's default highway build may be faster")
        line = {
            "metric": METRIC, "value": (1 if bands_mode else world) * mpix / (ms_per_step * 1e-3), "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if bands_mode else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": name, "frames_per_step": 1 if bands_mode else world,
                       "partition": "row bands of one frame, gathered on rank 0" if bands_mode else "one frame per rank", "out_mpix_per_frame": mpix,
                       "l2": "inputs_larger_than_l2 (321 MB source IR + 403 MB output per frame vs 126 MB L2)",
                       "texel_layout": "float4-padded" if args.padded == 1 else "interleaved-rgb (the library's rule for cubic jobs; bilinear configs below use 16-byte texels)",
                       "gather": "direct (L1)" if args.no_tiles else
                                 "footprint staged in shared memory by cp.async.bulk",
                       "arithmetic": args.arithmetic,  # eu_opts_t.reserved[1] & EU_OPT_CONTRACTED
                       "parity": parity_record("C2", args.arithmetic), "numa": numa,
                       "value_scope": "the render kernel alone (source staged and resident); cpu_baseline.value is the "
                                      "reference's whole payload() - compare `staging_plus_render` with it, or "
                                      "`value` with cpu_baseline.render_only_mpix_s"},
            "staging_plus_render": {"value": mpix / ((stage_ms + ms_per_step) * 1e-3), "unit": UNIT,
                                    "ms": stage_ms + ms_per_step, "scope": "device-timed staging + render = payload()'s scope"},
            "roofline": roof,
            "e2e": {"value": world * mpix / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h_src.numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4),
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "median_of_repetitions": {"value": world * mpix / (e2e_median_ms * 1e-3), "ms_per_step": e2e_median_ms,
                                              "repetitions_ms_per_step": [round(t / pipe_steps * 1e3, 3) for t in rep_times]},
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
            "configs": other, "configs_skipped": skipped, "run_s": time.perf_counter() - t_run0,
        }
        print(json.dumps(line))
    if hs is not None:
        eng.release(hs)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
# N > 1 (and --workload c5): BASELINE configs[4] split into row bands over the ranks
def c5_workload_name(scale, world):
    from envutil_b200 import c5
    (w, h), (W, H) = c5.sizes(scale)
    return ("C5: PTO stitch of 6 synthetic rectilinear %dx%d positions x 3 exposure brackets (hdr_merge per position, "
            "voronoi panorama) -> spherical %dx%d, float RGB, bilinear" % (w, h, W, H))


def ours_c5(args):
    import torch
    import torch.distributed as dist
    from envutil_b200 import bands as eu_bands, c5, synth
    from envutil_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = numa_setup(local)  # before any page-locked allocation: first touch places it next to this GPU
    synth.WORKERS = max(1, min(16, len(os.sched_getaffinity(0)) // max(1, min(world, 4))))
    t_run0 = time.perf_counter()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(v, op="max"):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    def every_rank(obj):
        if world == 1:
            return [obj]
        got = [None] * world
        dist.all_gather_object(got, obj)
        return got

    (w, h), (W, H) = c5.sizes(args.scale)
    eng = Engine(local)
    # ---- ONE frame in shared host memory; every rank page-locks and fills its own band ----------------
    tag = os.environ.get("MASTER_PORT", "0") + "_" + str(os.getppid() if world > 1 else os.getpid())
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    frame_path = os.path.join(shm, "eu_c5_frame_%s.f32" % tag)
    if rank == 0:
        np.memmap(frame_path, dtype=np.float32, mode="w+", shape=(H, W, 3)).flush()
    barrier()
    frame_np = np.memmap(frame_path, dtype=np.float32, mode="r+", shape=(H, W, 3))
    frame_t = torch.from_numpy(frame_np)
    t_setup = time.perf_counter()
    pl = c5.Pipeline(eng, torch, rank, world, args.scale, plan=args.c5_plan, contracted=args.arithmetic == "contracted",
                     host_frame=frame_t)
    band_t = frame_t[pl.row0:pl.row1]
    band_t.zero_()  # first touch by the rank that owns the band
    rt = torch.cuda.cudart()
    reg = rt.cudaHostRegister(band_t.data_ptr(), band_t.numel() * 4, 0)
    registered = int(reg) == 0 if not isinstance(reg, tuple) else int(reg[0]) == 0
    setup_s = time.perf_counter() - t_setup

    # ---- device-timed: stage A + brace + stage B on resident brackets, exactly K steps -------------------
    pl.upload()
    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        pl.step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.02)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = eng.launches
    tw0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        pl.step_device()
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    my_ms = ev0.elapsed_time(ev1) / args.steps
    ms_per_step = over_ranks(my_ms)
    launches = over_ranks(eng.launches - launches_before, "sum")
    clocks = None
    if sampler:
        time.sleep(0.15)
        clocks = ClockSampler.summarise(sampler.window(tw0, tw1))
        sampler.stop()
    # the stages on their own (same launches, timed separately; not part of `value`)
    a_ms = time_launches(torch, pl.stage_a, max(3, min(args.steps, 10)), 1)
    sb_ms = time_launches(torch, pl.stage_b_staging, max(3, min(args.steps, 10)), 1)
    b_ms = time_launches(torch, pl.stage_b, max(3, min(args.steps, 10)), 1)

    # ---- e2e: host rasters in, host frame out, every step ---------------------------------------------------
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 6))
    import gc
    gc.collect()
    gc.disable()
    for _ in range(2):
        pl.step_e2e()
    pl.finish()
    reps = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pl.step_e2e()
        pl.finish()
        barrier()
        reps.append(over_ranks(time.perf_counter() - t0))
    gc.enable()
    # the copies on their own: this rank's upload (H2D + brace) and its band's download
    up_ms = time_launches(torch, pl.upload, 2, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    band_t.copy_(pl.d_band[0], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    down_ms = e0.elapsed_time(e1)
    barrier()
    # ---- beside the timed regions: the bands gathered into rank 0's HBM over NCCL -----------------
    gather_ms = 0.0
    if world > 1:
        full = eu_bands.gather_ragged(pl.d_band[0], pl.bands, rank, dist)  # first use sets up the channels
        del full
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = eu_bands.gather_ragged(pl.d_band[0], pl.bands, rank, dist)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        dev_checksum = float(full[::61, ::67].double().sum().item()) if rank == 0 else None
        del full
    else:
        dev_checksum = float(pl.d_band[0][::61, ::67].double().sum().item())
    barrier()
    # ---- beside the timed regions: the SAME pipeline on one GPU of this box (rank 0 renders the whole panorama from the
    # whole brackets), so that a line of an N-GPU run carries its own single-GPU figure for this workload
    # (bench.py --gpus 1 measures configs[1], not configs[4])
    one_gpu = None
    if world > 1:
        if rank == 0:
            try:  # an extra beside the measurement: never at the price of the line (or of the ranks waiting at the barrier)
                p1 = c5.Pipeline(eng, torch, 0, 1, args.scale, plan=args.c5_plan, contracted=args.arithmetic == "contracted")
                p1.upload()
                ms1 = time_launches(torch, p1.step_device, max(3, min(args.steps, 10)), 3)
                p1.close()
                one_gpu = {"ms_per_step": ms1, "value": W * H / 1e6 / (ms1 * 1e-3), "unit": UNIT,
                           "what": "the whole panorama on rank 0's GPU alone, same code, same plan, brackets resident (device-timed)"}
            except Exception as e:  # noqa: BLE001
                one_gpu = {"error": "%s: %s" % (type(e).__name__, e)}
        barrier()
    per_rank = every_rank({"rank": rank, "rows": [pl.row0, pl.row1], "ms": my_ms, "stage_a_ms": a_ms, "stage_b_staging_ms": sb_ms,
                           "stage_b_ms": b_ms, "stage_a_mpix": c5.stage_a_pixels(pl.rects) * c5.POSITIONS / 1e6,
                           "stage_a_alg_bytes": pl.stage_a_alg_bytes(), "h2d_bytes": pl.h2d_bytes, "d2h_bytes": pl.d2h_bytes,
                           "upload_ms": up_ms, "download_ms": down_ms, "host_band_page_locked": registered, "numa": numa})
    if rank == 0:
        peak, peak_src = measured_peak()
        host_checksum = float(frame_np[::61, ::67].astype(np.float64).sum())
        slow = max(per_rank, key=lambda r: r["stage_a_ms"])
        roof = roofline_of(slow["stage_a_alg_bytes"], slow["stage_a_ms"], peak, peak_src, "C5A", args.arithmetic,
                           clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0, launches=0)
        roof["kernel"] = ("k_render<3,4,hdr_merge,...,6> (stage A: the slowest rank's launches over its rectangles; ncu figures "
                          "of the whole-raster launch are in the N=1 line's configs[C5A])")
        mpix = W * H / 1e6
        reps_sorted = sorted(reps)
        e2e_best, e2e_med = reps_sorted[0] / e2e_steps, reps_sorted[len(reps) // 2] / e2e_steps
        h2d_total = sum(r["h2d_bytes"] for r in per_rank)
        line = {
            "metric": METRIC, "value": mpix / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c5_workload_name(args.scale, world), "frames_per_step": 1, "out_mpix_per_frame": mpix,
                       "partition": "row bands of ONE panorama, one band per rank, sized by estimated cost (c5.plan_bands); "
                                    "reference work splitting: lines x segments of one frame (zimt/wielding.h:251-265)",
                       "c5": {"plan": args.c5_plan,
                              "what": ("stage A merges only the rectangles of every position that this rank's band of stage B "
                                       "samples (the voronoi winner is the position nearest in longitude: 49 % of each raster's "
                                       "columns; rows by the band's latitudes) - the panorama is bit-identical to the one from "
                                       "fully merged rasters (configs[] of the N=1 line: checksum_equals_full_merge)")
                              if args.c5_plan == "needed" else "stage A merges every texel of every position",
                              "bands": [list(b) for b in pl.bands]},
                       "l2": "inputs_larger_than_l2 (per rank: bracket rectangles + merged rasters + band >> 126 MB)",
                       "arithmetic": args.arithmetic,
                       "parity": {"C5A": parity_record("C5A", args.arithmetic), "C5B": parity_record("C5B", args.arithmetic)},
                       "value_scope": "stage A + brace + stage B, device-timed, brackets resident in HBM; max over ranks"},
            "roofline": roof,
            "e2e": {"value": mpix / e2e_best, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total),
                    "d2h_bytes_per_step": int(W * H * 12), "ms_per_step": e2e_best * 1e3, "steps": e2e_steps,
                    "median_of_repetitions": {"value": mpix / e2e_med, "ms_per_step": e2e_med * 1e3,
                                              "repetitions_ms_per_step": [round(t / e2e_steps * 1e3, 3) for t in reps]},
                    "scope": "every step, every rank: H2D of its rectangles of the 18 bracket rasters from page-locked host "
                             "memory (eu_source_write_rect) + brace + stage A + stage B + D2H of its band into ONE frame in "
                             "shared page-locked host memory over its own PCIe link; the download of step n overlaps the "
                             "upload of step n+1; wall clock, max over ranks, fastest of 3 repetitions",
                    "breakdown_ms": {"upload_max": max(r["upload_ms"] for r in per_rank),
                                     "download_max": max(r["download_ms"] for r in per_rank), "device_step": ms_per_step},
                    "host_frame_checksum": host_checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "multi_gpu": {"broadcast_ms": 0.0,
                          "source_distribution": "none: every rank uploads its own part of the sources (e2e.h2d_bytes_per_step is "
                                                 "the sum over ranks; per_rank[].upload_ms)",
                          "gather_ms": gather_ms,
                          "gather": "beside the timed regions: NCCL gather of the bands into rank 0's HBM; the e2e path needs none "
                                    "(every rank stores its band into the shared host frame, per_rank[].download_ms)",
                          "collectives_in_timed_region": 0, "nccl_ranks": world, "one_gpu_same_workload": one_gpu,
                          "per_rank": per_rank,
                          "checksum_device_gather": dev_checksum, "checksum_host_frame": host_checksum,
                          "frame_complete": dev_checksum == host_checksum},
            "cpu_baseline": None, "checksum": host_checksum, "setup_s": setup_s, "run_s": time.perf_counter() - t_run0,
        }
        print(json.dumps(line))
    barrier()
    if registered:
        rt.cudaHostUnregister(band_t.data_ptr())
    pl.close()
    eng.close()
    del frame_t, frame_np, band_t
    barrier()
    if rank == 0:
        try:
            os.unlink(frame_path)
        except OSError:
            pass
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_c5(args):
    """The unmodified reference on configs[4], on rank 0's host cores: a step = six `--synopsis hdr_merge --single 0` runs
    (one per position) + the panorama run over their outputs, wall clock minus the file shim's raster I/O."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from envutil_b200 import c5, euf, synth, workloads
    synth.WORKERS = max(1, min(16, os.cpu_count() or 1))
    exe = os.path.join(ROOT, "oracle", "_ref", "envutil_ref_fast")
    isa = "-march=x86-64-v3"
    try:
        flags = open("/proc/cpuinfo").read()
        if all((" " + f) in flags for f in ("avx512f", "avx512vl", "avx512bw", "avx512dq", "avx512cd")) and os.path.exists(exe + "512"):
            exe, isa = exe + "512", "-march=x86-64-v4"
    except OSError:
        pass
    (w, h), (W, H) = c5.sizes(args.scale)
    steps = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    mpix = W * H / 1e6
    ncores = os.cpu_count() or 1
    need = 18 * w * h * 12 + 6 * w * h * 12 + (64 << 20)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil_free("/dev/shm") > need else None
    with tempfile.TemporaryDirectory(prefix="eubench_c5_", dir=base) as d:
        fs = workloads.c5_facets(args.scale)
        paths = []
        for i, f in enumerate(fs):
            p = os.path.join(d, "bracket%02d.euf" % i)
            euf.write_euf(p, f.image)
            paths.append(p)
            f.image = None
        env_w = dict(os.environ)
        env_nw = dict(os.environ, EUSHIM_NOWRITE="1")

        def run(cmd, env):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError("reference failed: " + r.stderr[-1500:])
            rd = re.findall(r"eushim: read time ([0-9.eE+-]+) ms", r.stdout)
            wr = re.findall(r"eushim: write time ([0-9.eE+-]+) ms", r.stdout)
            return wall - ((float(rd[-1]) if rd else 0.0) + (float(wr[-1]) if wr else 0.0)) / 1e3

        from envutil_b200.job import FacetSpec
        ts = []
        for it in range(warm + steps):
            tot = 0.0
            merged = []
            for pz in range(c5.POSITIONS):
                sub = [FacetSpec(None, "rectilinear", c5.HFOV_DEG, yaw=c5.YAW_STEP_DEG * pz, eev=ev, width=w, height=h, nchannels=3)
                       for ev in c5.EVS]
                job, _ = workloads.c5_stage_a_geometry(sub, w, h)
                outp = os.path.join(d, "merged%d.euf" % pz)
                tot += run([exe, "-v"] + job.cli_args(paths[3 * pz:3 * pz + 3], outp), env_w)
                merged.append(outp)
            fsb = [FacetSpec(None, "rectilinear", c5.HFOV_DEG, yaw=c5.YAW_STEP_DEG * pz, width=w, height=h, nchannels=3)
                   for pz in range(c5.POSITIONS)]
            jobb, _ = workloads.c5_stage_b_geometry(fsb, args.scale)
            tot += run([exe, "-v"] + jobb.cli_args(merged, os.path.join(d, "pano.euf")), env_nw)
            if it >= warm:
                ts.append(tot)
    t = float(np.mean(ts))
    build = ("oracle/_ref/%s: unmodified reference sources, g++ -O3 %s, zimt goading back-end (highway / Vc are not in this "
             "image; the reference's default highway build may be faster)" % (os.path.basename(exe), isa))
    cb = {"value": mpix / t, "unit": UNIT, "cores": ncores, "threads": 2 * ncores, "kind": "reference", "build": build,
          "sample": "%d whole pipelines: 6 stage-A process runs + 1 stage-B run each = their payload()s, wall clock minus "
                    "raster file I/O" % steps}
    line = {"impl": "reference", "metric": METRIC, "value": mpix / t, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": c5_workload_name(args.scale, args.gpus), "frames_per_step": 1,
                       "note": "CPU reference runs once on rank 0's host cores; steps/warmup bounded to %d/%d pipelines"
                               % (steps, warm)},
            "cpu_baseline": cb, "e2e": {"value": mpix / t, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def shutil_free(path):
    import shutil
    try:
        return shutil.disk_usage(path).free
    except OSError:
        return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        wl = args.workload if args.workload != "auto" else ("c2" if max(world, args.gpus) == 1 else "c5")
        return reference_c5(args) if wl == "c5" else reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = args.workload if args.workload != "auto" else ("c2" if world == 1 else "c5")
    return ours_c5(args) if wl == "c5" else ours(args)


if __name__ == "__main__":
    sys.exit(main())
