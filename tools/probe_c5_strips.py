#!/usr/bin/env python3
"""One GPU plays rank R of W of the configs[4] pipeline: device time of a step with 2 / 4 / 8 column strips per position
(more strips follow the curved region more closely - fewer merged texels, more launches). Diagnostic tool."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from envutil_b200 import c5
from envutil_b200.engine import Engine

eng = Engine(0)


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for world, ranks in ((8, (0, 1, 3)), (4, (0, 1)), (1, (0,))):
    for rank in ranks:
        pl = c5.Pipeline(eng, torch, rank, world, 1, synth_inputs=False)
        row = {"world": world, "rank": rank}
        for k in (1, 2, 4, 8):
            pl.rects = c5.rects_for_band(pl.row0, pl.row1, pl.w, pl.h, pl.H, n_strips=k)
            row["strips%d" % k] = {"mpix": round(c5.stage_a_pixels(pl.rects) * 6 / 1e6, 2), "A_ms": round(timed(pl.stage_a), 4),
                                   "step_ms": round(timed(pl.step_device), 4)}
        pl.close()
        print(json.dumps(row), flush=True)
eng.close()
