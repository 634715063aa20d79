#!/usr/bin/env python3
"""Per-step wall-clock of the pipelined e2e leg (eu_source_upload_async / eu_render_async /
eu_job_wait) for the C2 workload: prints the time between consecutive eu_job_wait returns for a few
pipeline depths, next to the blocking pair. Diagnostic tool, not a bench."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from envutil_b200 import workloads
from envutil_b200.engine import Engine

job, alg = workloads.c2(1)
img = job.facets[0].image
h_src = torch.from_numpy(np.ascontiguousarray(img)).pin_memory()
eng = Engine(0)
st = job.structs(eng.lib)
t = st[0]
H, W, C = t.height, t.width, t.nchannels
ring = [torch.empty((H, W, C), dtype=torch.float32).pin_memory() for _ in range(4)]
for depth in (1, 2, 3, 4, 3, 3):
    pending, stamps = [], []
    t0 = time.perf_counter()
    for i in range(12):
        pending.append(eng.submit(job, st, [h_src.data_ptr()], ring[i % depth].data_ptr()))
        if len(pending) >= depth:
            eng.finish(pending.pop(0))
            stamps.append(time.perf_counter())
    while pending:
        eng.finish(pending.pop(0))
        stamps.append(time.perf_counter())
    d = np.diff(np.array([t0] + stamps)) * 1e3
    print("depth", depth, "total/step %.2f ms" % ((stamps[-1] - t0) * 1e3 / 12), " per-finish:", " ".join("%.1f" % x for x in d), flush=True)
eng.close()
