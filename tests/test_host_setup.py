"""Host-side set-up arithmetic of the C ABI (no GPU): the product's functions against the
independent restatement in the oracle and against known answers."""
import ctypes as C
import math

import numpy as np
import pytest

import harness
from envutil_b200 import capi

PRJ = range(7)


@pytest.mark.parametrize("prj", PRJ)
@pytest.mark.parametrize("w,h,hfov", [(1920, 1080, 90.0), (4096, 2048, 360.0), (600, 400, 47.3), (511, 777, 123.4)])
def test_extent_and_step_match_oracle(lib, prj, w, h, hfov):
    orc = harness.oracle()
    hf = math.radians(hfov)
    if prj in (capi.CUBEMAP, capi.BIATAN6):
        h = 6 * w
        hf = max(hf, math.pi / 2)
        hf = min(hf, math.radians(120.0))
    if prj == capi.RECTILINEAR:
        hf = min(hf, math.radians(150.0))
    a = (C.c_double * 4)()
    b = (C.c_double * 4)()
    lib.eu_get_extent(prj, w, h, hf, a)
    orc.orc_get_extent(prj, w, h, hf, b)
    assert list(a) == list(b)
    assert lib.eu_get_step(prj, w, h, hf) == orc.orc_get_step(prj, w, h, hf)


def test_setup_atan_is_the_contract_function(lib):
    """get_vfov / get_step use the double atan of include/eu_math.h (the pinned reference build
    interposes it for the whole program): 128x96 at hfov 90 is a case where glibc's atan gives a
    different last bit of the vertical extent, which showed as 1-ulp pixel differences."""
    orc = harness.oracle()
    a = (C.c_double * 4)()
    b = (C.c_double * 4)()
    for (w, h, hf) in ((128, 96, 90.0), (100, 75, 90.0), (6000, 4000, 100.0), (1920, 1080, 90.0)):
        lib.eu_get_extent(capi.RECTILINEAR, w, h, math.radians(hf), a)
        orc.orc_get_extent(capi.RECTILINEAR, w, h, math.radians(hf), b)
        assert list(a) == list(b)
        assert abs(a[3] - h / w * math.tan(math.radians(hf) / 2)) < 1e-15


def test_rotation_known_answer(lib):
    """roll 7, pitch -21, yaw 33 degrees, rows = images of e_x, e_y, e_z."""
    m = (C.c_double * 9)()
    lib.eu_rotation_matrix(math.radians(7), math.radians(-21), math.radians(33), 0, m)
    want = [0.808632611, 0.113774853, -0.577207598, -0.295934587, 0.926621656, -0.231937571,
            0.508464403, 0.358367967, 0.782966400]
    assert np.allclose(list(m), want, atol=2e-7)
    # README.md:971-976: yaw turns the camera right, pitch up (y points down), roll clockwise
    lib.eu_rotation_matrix(0, 0, math.radians(30), 0, m)
    assert np.allclose(list(m)[6:9], [0.5, 0, math.cos(math.radians(30))], atol=1e-6)
    lib.eu_rotation_matrix(0, math.radians(30), 0, 0, m)
    assert np.allclose(list(m)[6:9], [0, -0.5, math.cos(math.radians(30))], atol=1e-6)
    lib.eu_rotation_matrix(math.radians(30), 0, 0, 0, m)
    assert np.allclose(list(m)[0:3], [math.cos(math.radians(30)), 0.5, 0], atol=1e-6)


@pytest.mark.parametrize("angles", [(7, -21, 33), (0, 0, 0), (179, 89, -179), (-45.5, 12.25, 270)])
@pytest.mark.parametrize("inverse", [0, 1])
def test_rotation_matches_oracle_bitwise(lib, angles, inverse):
    orc = harness.oracle()
    r, p, y = (math.radians(v) for v in angles)
    a = (C.c_double * 9)()
    b = (C.c_double * 9)()
    lib.eu_rotation_matrix(r, p, y, inverse, a)
    orc.orc_rotation(r, p, y, inverse, b)
    assert list(a) == list(b)


def test_inverse_rotation_is_inverse(lib):
    a = (C.c_double * 9)()
    b = (C.c_double * 9)()
    lib.eu_rotation_matrix(0.3, -0.4, 1.2, 0, a)
    lib.eu_rotation_matrix(0.3, -0.4, 1.2, 1, b)
    A = np.array(list(a)).reshape(3, 3)
    B = np.array(list(b)).reshape(3, 3)
    assert np.allclose(A @ B, np.eye(3), atol=1e-6)


def test_facet_prepare_is_idempotent_and_scales_shift(lib):
    f = capi.Facet()
    f.projection, f.width, f.height, f.nchannels = capi.RECTILINEAR, 600, 400, 3
    f.hfov = math.radians(80)
    f.a, f.b, f.c, f.h, f.v = 0.01, -0.02, 0.03, 5.0, -3.0
    assert lib.eu_facet_prepare(C.byref(f)) == 0
    first = bytes(f)
    assert lib.eu_facet_prepare(C.byref(f)) == 0
    assert bytes(f) == first
    factor = abs(f.x1 - f.x0) / 600
    assert f.shift_h == 5.0 * factor and f.shift_v == -3.0 * factor
    assert f.has_lcp and f.has_shift and not f.has_shear
    assert f.d == 1.0 - (0.01 - 0.02 + 0.03)
    assert f.s == abs(f.y1 - f.y0) / 2.0  # the smaller half extent


def test_target_rules(lib):
    t = capi.Target()
    t.projection, t.width, t.hfov, t.nchannels = capi.CUBEMAP, 100, math.radians(90), 3
    assert lib.eu_target_prepare(C.byref(t)) == 0 and t.height == 600
    assert t.y0 == 6 * t.x0 and abs(t.x1 - 1.0) < 1e-15
    t = capi.Target()
    t.projection, t.width, t.hfov, t.nchannels = capi.SPHERICAL, 1001, math.radians(360), 3
    assert lib.eu_target_prepare(C.byref(t)) == 0 and (t.width, t.height) == (1002, 501)
    t = capi.Target()
    t.projection, t.width, t.hfov = capi.BIATAN6, 64, math.radians(60)
    assert lib.eu_target_prepare(C.byref(t)) == capi.EU_OK - 1  # EU_ERR_ARGUMENT: hfov < 90


@pytest.mark.parametrize("twine", [2, 3, 4, 7])
def test_box_spread(lib, twine):
    """make_spread (envutil_main.cc:1253-1355): twine^2 taps on a centred grid, equal weights."""
    t = capi.Target()
    t.projection, t.width, t.height, t.hfov, t.nchannels = capi.RECTILINEAR, 64, 64, 1.0, 3
    lib.eu_target_prepare(C.byref(t))
    f = capi.Facet()
    f.projection, f.width, f.height, f.nchannels, f.hfov = capi.SPHERICAL, 128, 64, 3, 2 * math.pi
    lib.eu_facet_prepare(C.byref(f))
    taps = (capi.Tap * 1024)()
    tw = C.c_int()
    o = capi.Opts()
    o.spline_degree, o.solo = 1, 0
    n = lib.eu_make_spread(C.byref(t), C.byref(o), 1, C.byref(f), twine, 1.0, 1.0, 0.0, 0.0, 8, taps, 1024, C.byref(tw))
    assert n == twine * twine and tw.value == twine
    xs = sorted({round(taps[i].x, 6) for i in range(n)})
    want = [-(twine - 1) / (2 * twine) + i / twine for i in range(twine)]
    assert np.allclose(xs, want, atol=1e-6)
    assert abs(sum(taps[i].w for i in range(n)) - 1.0) < 1e-5
    assert taps[1].x > taps[0].x and taps[0].y == taps[1].y  # x runs fastest


def test_auto_twine(lib):
    """twine_setup (envutil_main.cc:1450-1547): magnification < 1 -> twine = int(1 + 1/mag).
    C3a of BASELINE.json: lat/lon 16384 -> biatan6 4096 gives mag = 0.785 -> twine 2."""
    t = capi.Target()
    t.projection, t.width, t.hfov, t.nchannels = capi.BIATAN6, 4096, math.pi / 2, 3
    lib.eu_target_prepare(C.byref(t))
    f = capi.Facet()
    f.projection, f.width, f.height, f.nchannels, f.hfov = capi.SPHERICAL, 16384, 8192, 3, 2 * math.pi
    lib.eu_facet_prepare(C.byref(f))
    taps = (capi.Tap * 1024)()
    tw = C.c_int()
    o = capi.Opts()
    o.spline_degree, o.solo = 1, 0
    n = lib.eu_make_spread(C.byref(t), C.byref(o), 1, C.byref(f), -1, 1.0, 1.0, 0.0, 0.0, 8, taps, 1024, C.byref(tw))
    assert tw.value == 2 and n == 4


def test_cubemap_metrics(lib):
    """metrics_t (cubemap.h:233-400); SURVEY.md a-15: face 2048 -> section 2112, frame 32."""
    oi = (C.c_int32 * 4)()
    od = (C.c_double * 4)()
    assert lib.eu_cubemap_metrics(2048, math.pi / 2, 8, 64, oi, od) == 0
    assert list(oi) == [2112, 32, 32, 33]
    assert od[1] == 1024.0 and abs(od[0] - 1056 / 1024.0) < 1e-15
    assert lib.eu_cubemap_metrics(4096, math.pi / 2, 8, 64, oi, od) == 0
    assert oi[0] == 4160
    assert lib.eu_cubemap_metrics(64, math.pi / 2, 8, 48, oi, od) != 0  # tile size must be a power of two


def test_bspline_constants_match_oracle(lib):
    """The generated pole table (tools/gen_bspline_consts.py, mpmath) against the oracle's own
    derivation (Newton on the b-spline polynomial), and the cubic values the reference quotes
    (zimt/poles.h:1314)."""
    orc = harness.oracle()
    p = (C.c_longdouble * 4)()
    assert orc.orc_poles(3, p) == 1
    assert abs(float(p[0]) - (math.sqrt(3) - 2)) < 1e-15
    m = (C.c_float * 16)()
    orc.orc_weight_matrix(3, m)
    want = np.array([[1, 4, 1, 0], [-3, 0, 3, 0], [3, -6, 3, 0], [-1, 3, -3, 1]], dtype=np.float64) / 6.0
    assert np.allclose(np.array(list(m)).reshape(4, 4), want.astype(np.float32), atol=0)


def test_backend_option_bits(lib):
    """Job's back-end switches land in eu_opts_t.reserved as include/envutil_b200.h documents them; none is set
    by default (the measured, GPU-tested kernels are what a plain job runs)."""
    import copy
    import jobs
    job = copy.copy(jobs.JOBS["ll_rect_d3_rot"])
    o = job.structs(lib)[2]
    assert (o.reserved[0], o.reserved[1]) == (0, 0)
    job.padded = False  # 12-byte texels whatever the degree
    assert job.structs(lib)[2].reserved[0] == 2
    job.padded = None
    for field, word, bit in (("padded", 0, 1), ("no_tiles", 1, 1), ("no_spec", 1, 2), ("narrow_stores", 1, 4),
                             ("contracted", 1, 16)):
        j = copy.copy(job)
        setattr(j, field, True)
        o = j.structs(lib)[2]
        assert o.reserved[word] == bit and o.reserved[1 - word] == 0, field
