#!/usr/bin/env python3
"""Per-step wall-clock of the pipelined e2e leg (eu_source_upload_async / eu_render_async /
eu_job_wait) for the C2 workload: prints the time between consecutive eu_job_wait returns for a few
pipeline depths, next to the blocking pair. Diagnostic tool, not a bench.

  python tools/e2e_probe.py [--hold 1] [--blocking N] [--rows 1] [--n 40]
--hold: keep one staged source alive meanwhile; --blocking: N blocking upload+render pairs first;
--rows: some eu_render_rows launches on torch's stream first (what bench.py's device-timed leg does)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from envutil_b200 import capi, workloads
from envutil_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--hold", type=int, default=0)
ap.add_argument("--blocking", type=int, default=0)
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--n", type=int, default=12)
ap.add_argument("--depths", default="1,2,3,4,3,3")
a = ap.parse_args()
job, alg = workloads.c2(1)
img = job.facets[0].image
h_src = torch.from_numpy(np.ascontiguousarray(img)).pin_memory()
eng = Engine(0)
st = job.structs(eng.lib)
t, fa, o, taps, ntaps = st
H, W, Cc = t.height, t.width, t.nchannels
ring = [torch.empty((H, W, Cc), dtype=torch.float32).pin_memory() for _ in range(4)]
held = None
if a.hold or a.rows:
    d_src = h_src.cuda()
    held = eng.stage_device(job, [d_src.data_ptr()], st, stream=torch.cuda.current_stream().cuda_stream)
if a.rows:
    d_out = torch.empty((H, W, Cc), dtype=torch.float32, device="cuda")
    for _ in range(23):
        eng.render_rows(job, held, st, 0, H, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream, timed=False)
    torch.cuda.synchronize()
for _ in range(a.blocking):
    tmu, tmr = capi.Timing(), capi.Timing()
    h = capi.SourceH()
    capi.check(eng.lib.eu_source_upload(None, C.byref(fa[0]), C.byref(o), h_src.data_ptr(), C.byref(h), C.byref(tmu)), eng.lib)
    one = (capi.SourceH * 1)(h)
    capi.check(eng.lib.eu_render(C.byref(t), C.byref(o), 1, fa, one, taps, ntaps, ring[0].data_ptr(), C.byref(tmr)), eng.lib)
    capi.check(eng.lib.eu_source_release(h), eng.lib)
for depth in [int(v) for v in a.depths.split(",")]:
    pending, stamps = [], []
    t0 = time.perf_counter()
    for i in range(a.n):
        pending.append(eng.submit(job, st, [h_src.data_ptr()], ring[i % depth].data_ptr()))
        if len(pending) >= depth:
            eng.finish(pending.pop(0))
            stamps.append(time.perf_counter())
    while pending:
        eng.finish(pending.pop(0))
        stamps.append(time.perf_counter())
    d = np.diff(np.array([t0] + stamps)) * 1e3
    print("depth", depth, "total/step %.2f ms" % ((stamps[-1] - t0) * 1e3 / a.n), " per-finish:", " ".join("%.1f" % x for x in d), flush=True)
if held is not None:
    eng.release(held)
eng.close()
