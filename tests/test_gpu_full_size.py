"""Parity at BASELINE.json's full sizes: row bands of the full-size frame against the oracle
(bit-exact), plus size-independent properties (round trips, partition of unity)."""
import numpy as np
import pytest

import harness
from envutil_b200 import synth
from envutil_b200.job import FacetSpec, Job

pytestmark = pytest.mark.gpu


def _bands(h, n=3, rows=4):
    ys = np.linspace(0, h - rows, n).astype(int)
    return [(int(y), int(y) + rows) for y in ys]


def _check_bands(engine, job, n=3, rows=4):
    st = job.structs()
    hs = engine.stage(job, st)
    ohs = harness.oracle_sources(job, st)
    try:
        out = engine.render(job, sources=hs, structs=st)
        for r0, r1 in _bands(st[0].height, n, rows):
            ref = harness.oracle_render(job, rows=(r0, r1), sources=ohs)
            c = harness.compare(out[r0:r1], ref)
            assert c["n_diff"] == 0, (job.name, r0, c)
    finally:
        engine.release(hs)
        for oh in ohs:
            harness.oracle().orc_source_free(oh)
    return out


def test_c1_full_size(engine):
    """configs[0]: lat/lon 4096x2048 -> rectilinear 1920x1080 hfov 90, bilinear."""
    job = Job([FacetSpec(synth.latlon(4096), "spherical", 360.0)], "rectilinear", 90.0, 1920, 1080, name="C1")
    out = _check_bands(engine, job, n=5, rows=8)
    assert np.isfinite(out).all() and out.min() >= 0.0 and out.max() <= 1.0


def test_c2_full_size(engine):
    """configs[1]: cubemap 2048px faces -> spherical 8192x4096, cubic b-spline with prefilter."""
    job = Job([FacetSpec(synth.cubemap(2048), "cubemap", 90.0)], "spherical", 360.0, 8192, 4096, degree=3, name="C2")
    _check_bands(engine, job, n=4, rows=2)


def test_c4_reduced(engine):
    """configs[3] at half size: lat/lon 4096x2048 -> fisheye 2048^2 hfov 180, twine 4."""
    job = Job([FacetSpec(synth.latlon(4096), "spherical", 360.0)], "fisheye", 180.0, 2048, 2048, twine=4, name="C4/2")
    _check_bands(engine, job, n=3, rows=2)


def test_c3_round_trip(engine):
    """configs[2] at quarter size: lat/lon 4096 -> biatan6 1024px faces -> lat/lon; the GPU's
    round-trip error equals the oracle's (the pipelines are bit-identical) and is small."""
    ll = synth.latlon(4096, noise=0.0)
    fwd = Job([FacetSpec(ll, "spherical", 360.0)], "biatan6", 90.0, 1024, name="C3a/4")
    cube = engine.render(fwd)
    ref_rows = harness.oracle_render(fwd, rows=(3000, 3004))
    assert np.array_equal(cube[3000:3004], ref_rows)
    back = Job([FacetSpec(cube, "biatan6", 90.0)], "spherical", 360.0, 4096, 2048, name="C3b/4")
    rt = _check_bands(engine, back, n=3, rows=2)
    err = np.abs(rt.astype(np.float64) - ll)
    assert err.max() < 2e-2 and np.sqrt((err ** 2).mean()) < 1e-3, (err.max(), np.sqrt((err ** 2).mean()))


def test_constant_image_is_reproduced(engine):
    """Partition of unity: a constant source comes out constant (to float rounding) through the
    prefilter + cubic evaluation + twining, for every pixel that hits the source."""
    img = np.full((256, 512, 3), 0.625, dtype=np.float32)
    job = Job([FacetSpec(img, "spherical", 360.0)], "fisheye", 200.0, 300, 300, degree=3, twine=3, yaw=40.0, pitch=25.0)
    out = engine.render(job)
    # the full-sphere prefilter truncates its initial sums at tolerance 1e-4 (environment.h:385)
    assert np.abs(out - 0.625).max() < 1e-4


def test_c5_two_stage_reduced(engine):
    """configs[4] at 1/10 size, as the reference can run it: (A) per position,
    `--synopsis hdr_merge --single 0` of the three exposure brackets; (B) the voronoi panorama of
    the merged facets. Both stages bit-exact against the oracle; the merge reproduces the
    unclipped middle exposure where nothing is clipped."""
    from envutil_b200 import workloads
    fs = workloads.c5_facets(scale=10)
    merged, yaws = [], []
    for k in range(0, len(fs), 3):
        job, _ = workloads.c5_stage_a(fs[k:k + 3])
        out = engine.render(job)
        ref = harness.oracle_render(job)
        assert harness.compare(out, ref)["n_diff"] == 0, k
        merged.append(out)
        yaws.append(fs[k].yaw)
    mid = fs[0].image
    ok = mid < 0.2  # far from clipping in every bracket (the Eev-10 frame is 4x brighter)
    assert np.abs(merged[0] - mid)[ok].max() < 1e-5
    job, _ = workloads.c5_stage_b(merged, yaws, scale=10)
    out, idx = engine.render(job), engine.index_plane(job)
    ref, ridx = harness.oracle_render(job, want_index=True)
    assert harness.compare(out, ref)["n_diff"] == 0
    assert np.array_equal(idx, ridx)
    assert set(np.unique(idx)) == {-1, 0, 1, 2, 3, 4, 5}  # every facet wins somewhere; the poles are uncovered
