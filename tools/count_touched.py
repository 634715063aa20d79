#!/usr/bin/env python3
"""Algorithmic bytes of the BASELINE configs, counted exactly: output store + DISTINCT
source texels the job touches x 12 B. The distinct texels come from the oracle's own tap addresses
(oracle/eu_oracle.c: orc_touch_map marks every container texel a spline window reads), so the count
includes the brace texels a window reaches and nothing it does not. CPU only; C4 takes a minute.

  python tools/count_touched.py [--configs C1,C4] [--scale 1]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import harness  # noqa: E402
from envutil_b200 import workloads  # noqa: E402


def count(job):
    lib = harness.oracle()
    lib.orc_touch_map.restype = None
    lib.orc_touch_map.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_touch_map_size.restype = C.c_size_t
    lib.orc_touch_map_size.argtypes = [C.c_void_p]
    st = job.structs()
    t = st[0]
    hs = harness.oracle_sources(job, st)
    try:
        maps = []
        for h in hs:
            m = np.zeros(lib.orc_touch_map_size(h), dtype=np.uint8)
            lib.orc_touch_map(h, m.ctypes.data)
            maps.append(m)
        oh, ow = t.out_shape()
        out = np.empty((oh, ow, t.nchannels), dtype=np.float32)
        fa, o, taps, ntaps = st[1], st[2], st[3], st[4]
        rc = lib.orc_render(C.byref(t), C.byref(o), len(job.facets), fa, hs, taps, ntaps, 0, oh, out.ctypes.data, None, 0)
        assert rc == 0
        lib.orc_touch_map(None, None)
        touched = int(sum(int(m.sum()) for m in maps))
        n = int(sum(m.size for m in maps))
    finally:
        for h in hs:
            lib.orc_source_free(h)
    return touched, n, oh * ow


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C4")
    ap.add_argument("--scale", type=int, default=1)
    a = ap.parse_args()
    table = {"C1": workloads.c1, "C2": workloads.c2, "C3a": workloads.c3a, "C4": workloads.c4}
    for name in a.configs.split(","):
        if name == "C3b":    # the source's content does not matter for the addresses: any 1:6 raster
            from envutil_b200 import synth
            job, alg = workloads.c3b(synth.cubemap(4096 // a.scale))
        elif name == "C5A":  # one position: three brackets -> the geometry of the first
            job, alg = workloads.c5_stage_a(workloads.c5_facets(a.scale)[:3])
        elif name == "C5B":
            w, h = 6000 // a.scale, 4000 // a.scale
            merged = [np.zeros((h, w, 3), dtype=np.float32) for _ in range(6)]
            job, alg = workloads.c5_stage_b(merged, [60.0 * p for p in range(6)], scale=a.scale)
        else:
            job, alg = table[name](a.scale)
        touched, total, out_px = count(job)
        exact = out_px * 12 + touched * 12
        print(json.dumps({"config": name, "scale": a.scale, "out_px": out_px, "container_texels": total,
                          "touched_texels": touched, "touched_fraction": touched / total,
                          "algorithmic_bytes_exact": exact, "algorithmic_bytes_in_workloads": alg,
                          "bytes_per_px": exact / out_px}), flush=True)


if __name__ == "__main__":
    main()
