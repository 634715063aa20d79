"""The BASELINE.json configurations as Jobs, with their algorithmic byte counts.

Algorithmic bytes of one render = output store + distinct source texels touched x 12 B (the
compulsory HBM traffic); the per-config figures are derived in DESIGN.md. `scale` shrinks every
linear size by an integer factor for tests.
"""
import numpy as np

from . import synth
from .job import FacetSpec, Job

RGB = 12  # bytes per float RGB texel

# Distinct texels of the staged container(s) that the full-size job reads, counted from the tap addresses
# of the oracle (tools/count_touched.py; SURVEY.md 8d asks for the exact count). Other scales fall back to
# the fractions estimated in the survey.
EXACT_TOUCHED = {"C1": 627648, "C2": 25241100, "C3a": 120655786, "C3b": 100734138, "C4": 19806305,
                 "C5A": 72047136,   # one position: three brackets, every texel once
                 "C5B": 69790864}   # 48 % of the six merged images: only the winning facet is evaluated


def c1(scale=1, **kw):
    """configs[0]: lat/lon 4096x2048 -> rectilinear 1920x1080 hfov 90, bilinear, no twining."""
    src = synth.latlon(4096 // scale)
    job = Job([FacetSpec(src, "spherical", 360.0)], "rectilinear", 90.0, 1920 // scale, 1080 // scale, name="C1", **kw)
    out_px = job.width * job.height
    # 7.48 % of the source is inside the 90-degree view (counted numerically)
    alg = out_px * RGB + int(0.0748 * src.shape[0] * src.shape[1]) * RGB
    if scale == 1:
        alg = out_px * RGB + EXACT_TOUCHED["C1"] * RGB
    return job, alg


def c2(scale=1, **kw):
    """configs[1]: 1:6 cubemap, 2048px faces -> full spherical 8192x4096, cubic b-spline with
    prefilter. The whole sphere is seen, so all six faces (302 MB) are compulsory reads."""
    face = 2048 // scale
    src = synth.cubemap(face)
    job = Job([FacetSpec(src, "cubemap", 90.0)], "spherical", 360.0, 8192 // scale, 4096 // scale, degree=3,
              name="C2", **kw)
    alg = job.width * job.height * RGB + 6 * face * face * RGB
    if scale == 1:  # the cubic windows along the face edges reach into the support frame: +0.13 %
        alg = job.width * job.height * RGB + EXACT_TOUCHED["C2"] * RGB
    return job, alg


def c3a(scale=1, **kw):
    """configs[2], forward leg: lat/lon 16384x8192 -> biatan6 cubemap with 4096px faces."""
    src = synth.latlon(16384 // scale)
    job = Job([FacetSpec(src, "spherical", 360.0)], "biatan6", 90.0, 4096 // scale, degree=1, name="C3a", **kw)
    alg = job.width * 6 * job.width * RGB + src.shape[0] * src.shape[1] * RGB
    if scale == 1:  # 90 % of the source: where the cube faces sample coarser than the lat/lon grid, texels are skipped
        alg = job.width * 6 * job.width * RGB + EXACT_TOUCHED["C3a"] * RGB
    return job, alg


def c3b(cube, **kw):
    """configs[2], backward leg: that biatan6 cubemap -> lat/lon of the original size."""
    face = cube.shape[1]
    job = Job([FacetSpec(cube, "biatan6", 90.0)], "spherical", 360.0, 4 * face, 2 * face, degree=1, name="C3b", **kw)
    alg = job.width * job.height * RGB + 6 * face * face * RGB
    if face == 4096:
        alg = job.width * job.height * RGB + EXACT_TOUCHED["C3b"] * RGB
    return job, alg


def c4(scale=1, **kw):
    """configs[3]: lat/lon 8192x4096 -> fisheye 4096x4096 hfov 180, --twine 4 (16 sub-rays/px).
    A 180-degree fisheye sees the front hemisphere inside its image circle; the square's
    corners reach further: ~55.7 % of the source (numeric estimate)."""
    src = synth.latlon(8192 // scale)
    job = Job([FacetSpec(src, "spherical", 360.0)], "fisheye", 180.0, 4096 // scale, 4096 // scale, twine=4,
              name="C4", **kw)
    alg = job.width * job.height * RGB + int(0.557 * src.shape[0] * src.shape[1]) * RGB
    if scale == 1:
        alg = job.width * job.height * RGB + EXACT_TOUCHED["C4"] * RGB
    return job, alg


C5_BRACKETS = (12.0, 10.0, 14.0)  # Eev of the three exposures, middle exposure first


def c5_facets(scale=1, positions=6, brackets=C5_BRACKETS):
    """configs[4] inputs: `positions` rectilinear 6000x4000 views, yaw 60k degrees, hfov 100,
    each in three exposure brackets Eev 12/10/14 (images = clamp(scene * 2^(12-Eev), 0, 1))."""
    w, h = 6000 // scale, 4000 // scale
    fs = []
    for k in range(positions):
        yaw = 60.0 * k
        base = synth.rectilinear_facet(w, h, 100.0, yaw, 0.0, 0.0)
        for ev in brackets:
            img = np.clip(base * np.float32(2.0 ** (12.0 - ev)), 0.0, 1.0).astype(np.float32)
            fs.append(FacetSpec(img, "rectilinear", 100.0, yaw=yaw, eev=ev))
    return fs


def c5_stage_a(facets3, **kw):
    """C5 stage A: hdr_merge of the three brackets of one position into the geometry of the
    middle bracket (the reference cannot merge and stitch in one pass)."""
    f = facets3[0]
    w, h, _ = f.shape()
    # `--synopsis hdr_merge --single 0`: the target takes the geometry of the first (middle-exposure)
    # bracket, whose brighten is 1, so the un-brighten step of work() is a no-op
    job = Job(list(facets3), "rectilinear", f.hfov, w, h, yaw=f.yaw, synopsis="hdr_merge", single=0, name="C5A", **kw)
    alg = w * h * RGB * (1 + len(facets3))
    if (w, h, len(facets3)) == (6000, 4000, 3):
        alg = w * h * RGB + EXACT_TOUCHED["C5A"] * RGB
    return job, alg


def c5_stage_a_geometry(facets3, w, h, **kw):
    """Stage A for facets whose rasters live elsewhere (FacetSpec.image None, width/height set)."""
    return c5_stage_a(facets3, **kw)


def c5_stage_b(merged, yaws, hfov=100.0, scale=1, **kw):
    """C5 stage B: voronoi panorama of the merged facets -> spherical 16384x8192."""
    fs = [FacetSpec(m, "rectilinear", hfov, yaw=y) for m, y in zip(merged, yaws)]
    return c5_stage_b_geometry(fs, scale, **kw)


def c5_stage_b_geometry(fs, scale=1, **kw):
    job = Job(fs, "spherical", 360.0, 16384 // scale, 8192 // scale, name="C5B", **kw)
    alg = job.width * job.height * RGB + sum(f.shape()[0] * f.shape()[1] for f in fs) * RGB
    if scale == 1 and len(fs) == 6 and all(f.shape()[:2] == (6000, 4000) for f in fs):
        # only the winning facet of a pixel is evaluated: 48 % of the six merged images are ever read
        alg = job.width * job.height * RGB + EXACT_TOUCHED["C5B"] * RGB
    return job, alg
