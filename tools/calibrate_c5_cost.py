#!/usr/bin/env python3
"""Measure the per-pixel costs the band partition of envutil_b200/c5.py is built from (c5.row_costs: c_a, c_b, c_b0),
on one GPU: stage B on a polar band (no position in sight) and on an equatorial band, stage A per merged texel.

  python tools/calibrate_c5_cost.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from envutil_b200 import c5, synth  # noqa: E402
from envutil_b200.engine import Engine  # noqa: E402


def timed(fn, steps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    synth.WORKERS = max(1, min(16, os.cpu_count() or 1))
    eng = Engine(0)
    pl = c5.Pipeline(eng, torch, 0, 1, 1, plan="needed")
    pl.upload()
    pl.stage_a()
    pl.stage_b_staging()
    torch.cuda.synchronize()
    (w, h), (W, H) = c5.sizes(1)
    stream = torch.cuda.current_stream().cuda_stream
    out = pl.d_band[0]

    def band(r0, r1):
        return timed(lambda: eng.render_rows(pl.job_b, pl.hs_b, pl.st_b, r0, r1, out[r0].data_ptr(), stream, timed=False))
    rows = 1024
    polar = band(0, rows)
    equator = band(H // 2 - rows // 2, H // 2 + rows // 2)
    a_ms = timed(pl.stage_a)
    a_px = c5.stage_a_pixels(pl.rects) * c5.POSITIONS
    res = {"stage_b_polar_ps_per_px": polar * 1e9 / (rows * W), "stage_b_equator_ps_per_px": equator * 1e9 / (rows * W),
           "stage_a_ps_per_texel": a_ms * 1e9 / a_px, "stage_b_ms_full": band(0, H), "stage_a_ms": a_ms,
           "profile_ps_per_px": [band(r, r + 512) * 1e9 / (512 * W) for r in range(0, H, 512)]}
    print(json.dumps(res))
    pl.close()
    eng.close()


if __name__ == "__main__":
    main()
