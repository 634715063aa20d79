#!/usr/bin/env python3
"""Device-timed render Mpix/s and achieved algorithmic GB/s for every BASELINE.json config
(bench.py measures configs[1] only; this is the table in DESIGN.md section 6).

  python tools/bench_configs.py [--configs C1,C2,C3a,C3b,C4,C5] [--padded 0|1] [--steps 10] [--scale 1]

Each config: stage the source(s) (device-timed separately), 3 warm-up + `steps` timed launches of
the render kernel bracketed by CUDA events on the launching stream. Sources and outputs are all
larger than L2 except C1 (reported as such). One JSON line per config on stdout.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from envutil_b200 import workloads  # noqa: E402
from envutil_b200.engine import Engine  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def run(eng, job, alg, steps, padded, keep_output=False, flush=None):
    if isinstance(padded, (list, tuple)):  # several layouts on the same inputs: print all but the last here
        for p in padded[:-1]:
            print(json.dumps(run(eng, job, alg, steps, p, False, flush)[0]), flush=True)
        padded = padded[-1]
    job.padded = bool(padded & 1) if not (padded & 32) else None  # bit 5: the library's own rule
    job.no_tiles = bool(padded & 2)
    job.contracted = bool(padded & 8)
    st = job.structs(eng.lib)
    t = st[0]
    hs = eng.stage(job, st)
    stage_ms = sum(tm.render_ms for tm in eng.last_stage_timing)
    h2d_ms = sum(tm.h2d_ms for tm in eng.last_stage_timing)
    out = torch.empty((t.height, t.width, t.nchannels), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.render_rows(job, hs, st, 0, t.height, out.data_ptr(), stream, timed=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush is None:
        e0.record()
        for _ in range(steps):
            eng.render_rows(job, hs, st, 0, t.height, out.data_ptr(), stream, timed=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:  # small config: flush L2 between launches, time each launch on its own
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            e0.record()
            eng.render_rows(job, hs, st, 0, t.height, out.data_ptr(), stream, timed=False)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / steps
    eng.release(hs)
    mpix = t.width * t.height / 1e6
    gbs = alg / (ms * 1e-3) / 1e9
    res = {"config": job.name, "out": "%dx%d" % (t.width, t.height), "mpix": mpix, "ms": ms, "mpix_s": mpix / (ms * 1e-3),
           "alg_bytes": alg, "bytes_per_px": alg / (mpix * 1e6), "achieved_gbs": gbs, "frac_measured_peak": gbs / peak(),
           "frac_8tbs": gbs / 8000.0, "staging_ms": stage_ms, "h2d_ms": h2d_ms, "padded": int(padded),
           "l2": "flushed between launches" if flush is not None else "inputs larger than L2"}
    return res, (out.cpu().numpy() if keep_output else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3a,C3b,C4")
    ap.add_argument("--padded", default="0", help="comma list of variants: bit 0 = 16-byte texels, bit 1 = direct-gather kernel (no staging), bit 3 = contracted arithmetic (EU_OPT_CONTRACTED), bit 5 = texel layout by the library's rule")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--scale", type=int, default=1)
    a = ap.parse_args()
    a.padded = [int(v) for v in a.padded.split(",")]
    want = a.configs.split(",")
    from envutil_b200 import synth
    synth.WORKERS = max(1, min(16, os.cpu_count() or 1))
    eng = Engine(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    s = a.scale
    if "C1" in want:
        job, alg = workloads.c1(s)
        print(json.dumps(run(eng, job, alg, a.steps, a.padded, flush=flush)[0]), flush=True)
    if "C2" in want:
        job, alg = workloads.c2(s)
        print(json.dumps(run(eng, job, alg, a.steps, a.padded)[0]), flush=True)
    if "C3a" in want or "C3b" in want:
        job, alg = workloads.c3a(s)
        res, cube = run(eng, job, alg, max(3, a.steps // 2), a.padded, keep_output=True)
        print(json.dumps(res), flush=True)
        if "C3b" in want:
            ll = job.facets[0].image
            job2, alg2 = workloads.c3b(cube)
            res2, back = run(eng, job2, alg2, max(3, a.steps // 2), a.padded, keep_output=True)
            err = np.abs(back.astype(np.float64) - ll)
            res2["round_trip_max_abs"] = float(err.max())
            res2["round_trip_rms"] = float(np.sqrt((err ** 2).mean()))
            print(json.dumps(res2), flush=True)
            del back, err
        del cube
    if "C4" in want:
        job, alg = workloads.c4(s)
        print(json.dumps(run(eng, job, alg, max(3, a.steps // 2), a.padded)[0]), flush=True)
    if "C5" in want:
        t0 = time.time()
        fs = workloads.c5_facets(s)
        merged, yaws = [], []
        a_ms, a_alg, a_mpix = 0.0, 0, 0.0
        for k in range(0, len(fs), 3):
            job, alg = workloads.c5_stage_a(fs[k:k + 3])
            res, m = run(eng, job, alg, 3, a.padded, keep_output=True)
            merged.append(m)
            yaws.append(fs[k].yaw)
            a_ms += res["ms"]
            a_alg += alg
            a_mpix += res["mpix"]
        gbs = a_alg / (a_ms * 1e-3) / 1e9
        print(json.dumps({"config": "C5A (6 x hdr_merge of 3 brackets)", "mpix": a_mpix, "ms": a_ms,
                          "mpix_s": a_mpix / (a_ms * 1e-3), "alg_bytes": a_alg, "achieved_gbs": gbs,
                          "frac_measured_peak": gbs / peak(), "padded": a.padded[-1]}), flush=True)
        job, alg = workloads.c5_stage_b(merged, yaws, scale=s)
        res, _ = run(eng, job, alg, 3, a.padded)
        res["synth_s"] = time.time() - t0
        print(json.dumps(res), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
