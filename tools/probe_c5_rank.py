#!/usr/bin/env python3
"""One GPU plays rank R of a W-rank configs[4] job (envutil_b200/c5.py): device time of stage A and of the whole step
with the merges on one stream and dealt to several. Diagnostic tool (inputs are not uploaded: timing only)."""
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


for world, ranks in ((8, (0, 1, 3)), (4, (0, 1)), (2, (0,)), (1, (0,))):
    for rank in ranks:
        row = {"world": world, "rank": rank}
        for ns in (0, 2, 4):
            pl = c5.Pipeline(eng, torch, rank, world, 1, synth_inputs=False, a_streams=ns)
            row["rects"] = [list(r) for r in pl.rects]
            row["A_ms_streams%d" % ns] = round(timed(pl.stage_a), 4)
            row["step_ms_streams%d" % ns] = round(timed(pl.step_device), 4)
            pl.close()
        print(json.dumps(row), flush=True)
eng.close()
