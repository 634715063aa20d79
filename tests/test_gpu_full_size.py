"""Parity at BASELINE.json's FULL sizes, whole frames, every config: the CUDA path - both arithmetics of the library -
against the unmodified reference binaries under oracle/_ref/ run on this box's host cores (tools/full_parity.py).

  exact arithmetic       0 differing floats against the pinned-math build (envutil_ref_pm) on every config
  contracted arithmetic  max relative difference <= 1e-5 against the pinned-math build (tolerance of BASELINE.json's
                         north_star; eps 1e-3 in the denominator), hdr_merge's near-zero denominators <= 5e-5
  both                   max / RMS against the stock-libm build are PRINTED, with the tie-band pixel count: two builds
                         of the reference differ from each other by the same amount (ref_self)
plus size-independent properties (partition of unity, the C3 round trip) and the configs[4] pipeline of
envutil_b200/c5.py (row bands, rectangles of the position rasters) against the two-stage oracle at 1/10 size.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import pytest

import harness
from envutil_b200 import synth
from envutil_b200.job import FacetSpec, Job

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

CONFIGS = ["C1", "C2", "C3a", "C3b", "C4", "C5A", "C5B"]


@pytest.fixture(scope="module")
def records(engine):
    """All configs once (inputs are shared: C3b reads C3a's frame, C5B the merged rasters of C5A)."""
    import full_parity
    from envutil_b200 import synth as sy
    if not (harness.ref_binary("pm") and harness.ref_binary("libm")):
        pytest.skip("oracle/_ref is not built")
    sy.WORKERS = max(1, min(16, os.cpu_count() or 1))
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    workdir = tempfile.mkdtemp(prefix="euparity_", dir=base)
    out = {}
    try:
        for both in full_parity.configs_iter(engine, CONFIGS, workdir, 1, True, None):
            name = both["exact"]["config"]
            out[name] = both
            slim = {ar: {k: v for k, v in both[ar].items() if k != "positions"} for ar in both}
            print("PARITY", json.dumps(slim), flush=True)
    finally:
        sy.WORKERS = 1
        shutil.rmtree(workdir, ignore_errors=True)
    dst = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dst):  # keep the numbers of this run (copied to profiles/ by hand)
        for ar in ("exact", "contracted"):
            json.dump({"arithmetic": ar, "scale": 1, "eps": full_parity.EPS, "tolerance": full_parity.TOL,
                       "cores": os.cpu_count(), "configs": [out[c][ar] for c in CONFIGS if c in out],
                       "from": "tests/test_gpu_full_size.py"},
                      open(os.path.join(dst, "parity_full_from_tests_%s.json" % ar), "w"), indent=1)
    return out


@pytest.mark.parametrize("name", CONFIGS)
def test_full_size_whole_frame(records, name):
    both = records[name]
    ex, co = both["exact"], both["contracted"]
    assert ex["vs_pinned"]["n_diff"] == 0, (name, ex["vs_pinned"])
    bar = 5e-5 if name == "C5A" else 1e-5  # hdr_merge divides by a sum of weights that can be close to zero
    assert co["vs_pinned"]["max_rel"] <= bar, (name, co["vs_pinned"])
    assert co["vs_pinned"]["rms_rel"] <= 5e-7, (name, co["vs_pinned"])
    if "vs_libm" in ex:  # how far the libm build is: no further than the reference's two builds are apart
        assert ex["vs_libm"]["max_rel"] == pytest.approx(ex["ref_self"]["max_rel"]), name
        assert ex["vs_libm"]["rms_rel"] <= 5e-5, (name, ex["vs_libm"])
    if name == "C3b":  # round trip lat/lon -> biatan6 -> lat/lon: the GPU's error is the reference's
        assert ex["round_trip"]["gpu"] == ex["round_trip"]["reference"]
        assert ex["round_trip"]["gpu"]["rms"] < 0.05


def test_constant_image_is_reproduced(engine):
    """Partition of unity: a constant source comes out constant (to float rounding) through the
    prefilter + cubic evaluation + twining, for every pixel that hits the source."""
    img = np.full((256, 512, 3), 0.625, dtype=np.float32)
    job = Job([FacetSpec(img, "spherical", 360.0)], "fisheye", 200.0, 300, 300, degree=3, twine=3, yaw=40.0, pitch=25.0)
    out = engine.render(job)
    # the full-sphere prefilter truncates its initial sums at tolerance 1e-4 (environment.h:385)
    assert np.abs(out - 0.625).max() < 1e-4


import functools


@functools.lru_cache(maxsize=2)
def _two_stage_oracle(scale):
    from envutil_b200 import workloads
    fs = workloads.c5_facets(scale=scale)
    merged, yaws = [], []
    for k in range(0, len(fs), 3):
        job, _ = workloads.c5_stage_a(fs[k:k + 3])
        merged.append(harness.oracle_render(job))
        yaws.append(fs[k].yaw)
    job, _ = workloads.c5_stage_b(merged, yaws, scale=scale)
    return fs, merged, harness.oracle_render(job)


def test_c5_two_stage_reduced(engine):
    """configs[4] at 1/10 size, as the reference can run it: (A) per position,
    `--synopsis hdr_merge --single 0` of the three exposure brackets; (B) the voronoi panorama of
    the merged facets. Both stages bit-exact against the oracle; the merge reproduces the
    unclipped middle exposure where nothing is clipped."""
    from envutil_b200 import workloads
    fs, omerged, opano = _two_stage_oracle(10)
    merged, yaws = [], []
    for k in range(0, len(fs), 3):
        job, _ = workloads.c5_stage_a(fs[k:k + 3])
        out = engine.render(job)
        assert harness.compare(out, omerged[k // 3])["n_diff"] == 0, k
        merged.append(out)
        yaws.append(fs[k].yaw)
    mid = fs[0].image
    ok = mid < 0.2  # far from clipping in every bracket (the Eev-10 frame is 4x brighter)
    assert np.abs(merged[0] - mid)[ok].max() < 1e-5
    job, _ = workloads.c5_stage_b(merged, yaws, scale=10)
    out, idx = engine.render(job), engine.index_plane(job)
    ref, ridx = harness.oracle_render(job, want_index=True)
    assert harness.compare(out, opano)["n_diff"] == 0
    assert np.array_equal(idx, ridx)
    assert set(np.unique(idx)) == {-1, 0, 1, 2, 3, 4, 5}  # every facet wins somewhere; the poles are uncovered


@pytest.mark.parametrize("world,plan", [(1, "needed"), (3, "needed"), (8, "needed"), (1, "full")])
def test_c5_pipeline_bands_equal_two_stage_oracle(engine, world, plan):
    """envutil_b200/c5.py at 1/10 size, the ranks of a `world`-GPU job played one after the other on this GPU: every
    rank uploads its rectangles of the bracket rasters, merges them in place, stitches its band and stores it
    into the one host frame - which equals the oracle's two-stage panorama bit for bit (two e2e steps each:
    the second one runs with the band buffers swapped)."""
    import torch
    from envutil_b200 import c5
    _, _, opano = _two_stage_oracle(10)
    (w, h), (W, H) = c5.sizes(10)
    frame = torch.full((H, W, 3), -3.0, dtype=torch.float32).pin_memory()
    for rank in range(world):
        pl = c5.Pipeline(engine, torch, rank, world, 10, plan=plan, host_frame=frame)
        try:
            pl.step_e2e()
            pl.step_e2e()
            pl.finish()
        finally:
            pl.close()
    assert harness.compare(frame.numpy(), opano)["n_diff"] == 0
