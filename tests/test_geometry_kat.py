"""Known-answer checks restated from the reference's geometry.cc, run through the whole path
(CPU: the oracle; GPU: the kernels):

(a) ray <-> plane round trips (geometry.cc:283-331): X_to_ray followed by ray_to_X is the identity for every
    projection. Here: a source in projection X rendered into a target of the same projection, field of view and
    size with no rotation - every target pixel's ray must land on the centre of the source texel with the same
    index, so the output reproduces the input (bilinear interpolation at a texel centre returns the texel).
(b) stepper == linspace + functor (geometry.cc:461-984): the incremental steppers produce the rays the plain
    projection functors produce for the same planar coordinates. Here: the same job through the projection's own
    stepper and through the generic stepper (forced by a translation too small to move any ray measurably).
Tolerances are those of float coordinates on a smooth scene: |error| <= gradient x a few 1e-4 texels.
"""
import copy

import numpy as np
import pytest

import harness
from envutil_b200 import synth
from envutil_b200.job import FacetSpec, Job


def _smooth(h, w):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([0.5 + 0.3 * np.sin(x / w * 5.0) * np.cos(y / h * 3.0), 0.4 + 0.4 * (x / w) * (y / h),
                    0.6 - 0.2 * np.cos((x + 2 * y) / (w + h) * 7.0)], axis=2)
    return np.ascontiguousarray(img, dtype=np.float32)


ROUND_TRIPS = {
    # projection: (hfov, width, height)
    "spherical": (360.0, 128, 64),
    "spherical_partial": (120.0, 96, 48),
    "cylindrical": (200.0, 120, 60),
    "rectilinear": (90.0, 96, 64),
    "stereographic": (150.0, 80, 80),
    "fisheye": (170.0, 80, 80),
    "cubemap": (90.0, 24, 144),
    "biatan6": (90.0, 24, 144),
}


def _round_trip_job(name):
    hfov, w, h = ROUND_TRIPS[name]
    prj = name.split("_")[0]
    img = _smooth(h, w)
    if prj in ("cubemap", "biatan6"):
        return Job([FacetSpec(img, prj, hfov)], prj, hfov, w, name="rt_" + name), img
    return Job([FacetSpec(img, prj, hfov)], prj, hfov, w, h, name="rt_" + name), img


def _check_round_trip(render, name):
    job, img = _round_trip_job(name)
    out = render(job)
    assert out.shape == img.shape
    err = np.abs(out.astype(np.float64) - img)
    # a texel-centre hit reproduces the texel up to the float error of the coordinate times the local gradient
    assert err.max() < 2e-4, (name, float(err.max()))


@pytest.mark.parametrize("name", sorted(ROUND_TRIPS))
def test_ray_plane_round_trip_oracle(name):
    _check_round_trip(harness.oracle_render, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ROUND_TRIPS))
def test_ray_plane_round_trip_gpu(engine, name):
    _check_round_trip(engine.render, name)


# PanoTools' translation model only covers rays in front of the translation plane (tf3d_t marks the others
# (0, 0, -inf), geometry.h:1886-1925), so the views stay inside that half-space: 100 degrees around yaw 10, and of the
# cubemap targets the FRONT face (rows 4w .. 5w)
STEPPER_TARGETS = [("spherical", 100.0, 160, 80), ("cylindrical", 100.0, 150, 50), ("rectilinear", 100.0, 120, 80),
                   ("stereographic", 100.0, 90, 90), ("fisheye", 100.0, 90, 90), ("cubemap", 90.0, 40, 0),
                   ("biatan6", 90.0, 40, 0)]


def _stepper_jobs(prj, hfov, w, h):
    src = synth.latlon(256, noise=0.0)
    fast = Job([FacetSpec(src, "spherical", 360.0)], prj, hfov, w, h, yaw=10.0, pitch=-10.0, roll=6.0)
    slow = copy.deepcopy(fast)
    slow.facets[0].tr_x = 1e-9  # has_translation: the generic stepper (linspace + X_to_ray functor + tf3d_t)
    return fast, slow


def _check_stepper(render, prj, hfov, w, h):
    fast, slow = _stepper_jobs(prj, hfov, w, h)
    a, b = render(fast), render(slow)
    assert a.shape == b.shape
    if prj in ("cubemap", "biatan6"):
        a, b = a[4 * w:5 * w], b[4 * w:5 * w]
    d = np.abs(a.astype(np.float64) - b)
    # the generic stepper shifts by 1e-9 and rounds differently: a few ulp of the ray, times the scene's gradient
    assert d.max() < 2e-4, (prj, float(d.max()))
    assert (a != b).any() or prj == "rectilinear"  # ... and it IS another code path


@pytest.mark.parametrize("prj,hfov,w,h", STEPPER_TARGETS)
def test_stepper_equals_functor_oracle(prj, hfov, w, h):
    _check_stepper(harness.oracle_render, prj, hfov, w, h)


@pytest.mark.gpu
@pytest.mark.parametrize("prj,hfov,w,h", STEPPER_TARGETS)
def test_stepper_equals_functor_gpu(engine, prj, hfov, w, h):
    _check_stepper(engine.render, prj, hfov, w, h)
