#!/usr/bin/env python3
"""Stage one source (default: the C2 cubemap, cubic) twice; used under `ncu --metrics
gpu__time_duration.sum` to list the staging kernels. Random texels (values do not matter)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from envutil_b200.engine import Engine
from envutil_b200.job import FacetSpec, Job

kind = sys.argv[1] if len(sys.argv) > 1 else "cubemap"
rng = np.random.default_rng(1)
if kind == "cubemap":
    img = rng.random((6 * 2048, 2048, 3), dtype=np.float32)
    job = Job([FacetSpec(img, "cubemap", 90.0)], "spherical", 360.0, 1024, 512, degree=3)
else:
    img = rng.random((4096, 8192, 3), dtype=np.float32)
    job = Job([FacetSpec(img, "spherical", 360.0)], "rectilinear", 90.0, 1024, 512, degree=3)
eng = Engine(0)
for _ in range(2):
    hs = eng.stage(job)
    print("staging ms", eng.last_stage_timing[0].render_ms, "h2d", eng.last_stage_timing[0].h2d_ms)
    eng.release(hs)
eng.close()
