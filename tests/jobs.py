"""The small parity jobs: every stepper x source kind x interpolator x synopsis combination of
the hot path at sizes the CPU oracle and the reference binary finish in
well under a second. Shared by the oracle-vs-reference tests, the golden-fixture generator
(tools/make_golden.py) and the GPU parity tests. Sources are deterministic (envutil_b200.synth),
so a job is fully described by its name.
"""
import functools

import numpy as np

from envutil_b200 import synth
from envutil_b200.job import FacetSpec, Job


@functools.lru_cache(maxsize=None)
def _ll(w, h=None):
    return synth.latlon(w, h)


@functools.lru_cache(maxsize=None)
def _cm(face, biatan6=False, hfov=90.0):
    return synth.cubemap(face, biatan6=biatan6, hfov_deg=hfov)


@functools.lru_cache(maxsize=None)
def _rect(w, h, hfov, yaw=0.0, pitch=0.0, roll=0.0, gain=1.0):
    return synth.rectilinear_facet(w, h, hfov, yaw, pitch, roll, gain=gain)


def _grey(img):
    return np.ascontiguousarray(img[:, :, 1:2])


def _alpha(h, w, seed):
    """A smooth alpha plane in [0,1] with fully opaque and fully transparent regions."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    r = np.hypot((x - w * (0.45 + 0.1 * seed)) / w, (y - h * 0.5) / h)
    return np.clip(1.6 - 3.2 * r, 0.0, 1.0).astype(np.float32)


def _rgba(img, seed=0):
    """Associated-alpha RGBA (colour premultiplied), as OIIO hands it over."""
    a = _alpha(img.shape[0], img.shape[1], seed)[:, :, None]
    return np.ascontiguousarray(np.concatenate([img * a, a], axis=2).astype(np.float32))


def _ga(img, seed=0):
    a = _alpha(img.shape[0], img.shape[1], seed)[:, :, None]
    return np.ascontiguousarray(np.concatenate([img[:, :, 1:2] * a, a], axis=2).astype(np.float32))


def _ll_facet(w, **kw):
    return FacetSpec(_ll(w), "spherical", 360.0, **kw)


ROT = dict(yaw=33.0, pitch=-21.0, roll=7.0)


def _voronoi_facets(w=96, h=64, hfov=80.0, n=4, step=70.0, pitch=(0, 12, -9, 5, -14, 8), roll=(0, 3, -5, 2, 0, -4)):
    fs = []
    for k in range(n):
        y, p, r = step * k - 100.0, float(pitch[k % 6]), float(roll[k % 6])
        fs.append(FacetSpec(_rect(w, h, hfov, y, p, r), "rectilinear", hfov, yaw=y, pitch=p, roll=r))
    return fs


def _bracket_facets(w=96, h=64, hfov=70.0, yaw=20.0):
    # Eev 12 / 10 / 14 (mean 12): brighten 1, 1/4, 4; bracket images = clamp(f * 2^(12-Eev), 0, 1),
    # so the Eev-10 frame clips and exercises the over-exposure logic
    fs = []
    for ev in (12.0, 10.0, 14.0):
        gain = 2.0 ** (12.0 - ev)
        fs.append(FacetSpec(_rect(w, h, hfov, yaw, 0.0, 0.0, gain), "rectilinear", hfov, yaw=yaw, eev=ev))
    return fs


def _lens_facets():
    # PanoTools lens polynomial + shift + shear on the source side (environment.h:240-284)
    return [FacetSpec(_rect(96, 64, 75.0, -30.0, 5.0, 0.0), "rectilinear", 75.0, yaw=-30.0, pitch=5.0,
                      a=0.02, b=-0.05, c=0.01, d=3.5, e=-2.25, g=1.5, t=-0.75),
            FacetSpec(_rect(96, 64, 75.0, 25.0, -8.0, 3.0), "rectilinear", 75.0, yaw=25.0, pitch=-8.0, roll=3.0,
                      a=-0.01, b=0.03, c=0.0),
            FacetSpec(_ll(96, 96), "fisheye", 140.0, yaw=90.0, b=-0.02, d=-1.0)]


def _translated_facets():
    # PanoTools translation (TrX/TrY/TrZ, optionally on a tilted plane Tpy/Tpp): generic stepper path
    return [FacetSpec(_rect(96, 64, 80.0, -30.0, 5.0, 0.0), "rectilinear", 80.0, yaw=-30.0, pitch=5.0,
                      tr_x=0.05, tr_y=-0.03, tr_z=0.02),
            FacetSpec(_rect(96, 64, 80.0, 30.0, -4.0, 2.0), "rectilinear", 80.0, yaw=30.0, pitch=-4.0, roll=2.0,
                      tr_x=-0.04, tr_y=0.02, tr_z=0.1, tp_y=10.0, tp_p=-5.0),
            FacetSpec(_rect(96, 64, 80.0, 100.0, 0.0, 0.0), "rectilinear", 80.0, yaw=100.0)]


def _build():
    J = {}

    def add(name, job):
        job.name = name
        J[name] = job

    # --- lat/lon source, all seven target projections -----------------------------------
    add("ll_rect_d1", Job([_ll_facet(256)], "rectilinear", 90.0, 96, 54))                       # C1 shape
    add("ll_rect_d1_rot", Job([_ll_facet(256)], "rectilinear", 90.0, 96, 54, **ROT))
    add("ll_rect_d3_rot", Job([_ll_facet(256)], "rectilinear", 70.0, 96, 54, degree=3, **ROT))
    # angles that are not float-representable: the target's go through float, the facet's through double
    add("ll_rect_d1_oddangles", Job([_ll_facet(256, yaw=12.34, pitch=-3.21, roll=0.77)], "rectilinear", 47.3, 96, 54,
                                    yaw=33.3, pitch=-21.7, roll=7.1))
    add("ll_rect_d2", Job([_ll_facet(128)], "rectilinear", 100.0, 80, 60, degree=2, yaw=170.0))
    add("ll_rect_d5", Job([_ll_facet(128)], "rectilinear", 60.0, 64, 64, degree=5, pitch=80.0))
    add("ll_rect_d7", Job([_ll_facet(128)], "rectilinear", 60.0, 40, 40, degree=7, pitch=-88.0, roll=45.0))
    add("ll_rect_d0", Job([_ll_facet(128)], "rectilinear", 60.0, 40, 40, degree=0, yaw=-120.0))
    add("ll_sph_d1", Job([_ll_facet(256)], "spherical", 360.0, 128, 64))
    add("ll_sph_d3_rot_wide", Job([_ll_facet(128)], "spherical", 360.0, 1100, 8, degree=3, **ROT))  # > 2 segments
    add("ll_cyl_d1", Job([_ll_facet(256)], "cylindrical", 200.0, 128, 64, pitch=10.0))
    add("ll_cyl_d1_tw2", Job([_ll_facet(256)], "cylindrical", 360.0, 600, 40, twine=2, **ROT))     # normalised cyl.
    add("ll_ster_d1", Job([_ll_facet(256)], "stereographic", 200.0, 96, 96, yaw=-45.0))
    add("ll_fish_d1_tw4", Job([_ll_facet(256)], "fisheye", 180.0, 96, 96, twine=4))              # C4 shape
    add("ll_fish_d3_rot", Job([_ll_facet(128)], "fisheye", 220.0, 80, 80, degree=3, **ROT))
    add("ll_cube_d1", Job([_ll_facet(256)], "cubemap", 90.0, 48))
    add("ll_cube_d3_rot", Job([_ll_facet(128)], "cubemap", 100.0, 32, degree=3, **ROT))
    add("ll_ba6_d1", Job([_ll_facet(256)], "biatan6", 90.0, 48))                                  # C3a shape
    add("ll_ba6_d1_tw2", Job([_ll_facet(256)], "biatan6", 90.0, 40, twine=2, yaw=12.0))
    # twining variants: gaussian weights with threshold, wide kernel
    add("ll_rect_d1_tw3_sigma", Job([_ll_facet(256)], "rectilinear", 90.0, 64, 36, twine=3, twine_width=1.5,
                                    twine_sigma=1.2, twine_threshold=0.05))
    add("ll_rect_d3_tw2", Job([_ll_facet(128)], "rectilinear", 50.0, 64, 36, degree=3, twine=2, **ROT))
    # non-2:1 and partial lat/lon sources (plain prefilter, REFLECT / PERIODIC without the pole brace)
    add("llpart_rect_d3", Job([FacetSpec(_ll(128, 64)[8:56, 16:112].copy(), "spherical", 270.0)], "rectilinear", 80.0,
                              64, 48, degree=3, yaw=10.0))
    add("ll360x90_rect_d3", Job([FacetSpec(_ll(128, 64)[16:48].copy(), "spherical", 360.0)], "rectilinear", 70.0,
                                64, 40, degree=3, yaw=175.0))
    add("cyl360_src_sph_d2", Job([FacetSpec(_ll(128, 64)[8:56].copy(), "cylindrical", 360.0)], "spherical", 360.0,
                                 96, 48, degree=2))
    # --- cubemap / biatan6 sources ------------------------------------------------------
    add("cm_sph_d1", Job([FacetSpec(_cm(64), "cubemap", 90.0)], "spherical", 360.0, 192, 96))
    add("cm_sph_d3", Job([FacetSpec(_cm(64), "cubemap", 90.0)], "spherical", 360.0, 192, 96, degree=3))  # C2 shape
    add("cm_sph_d3_rot", Job([FacetSpec(_cm(48), "cubemap", 90.0)], "spherical", 360.0, 160, 80, degree=3, **ROT))
    add("cm_rect_d1_tw3_rot", Job([FacetSpec(_cm(64), "cubemap", 90.0)], "rectilinear", 110.0, 80, 60, twine=3, **ROT))
    add("cm100_sph_d2", Job([FacetSpec(_cm(64, hfov=100.0), "cubemap", 100.0)], "spherical", 360.0, 128, 64, degree=2))
    add("cm_sph_d5", Job([FacetSpec(_cm(40), "cubemap", 90.0)], "spherical", 360.0, 96, 48, degree=5))
    add("ba6_sph_d1", Job([FacetSpec(_cm(64, True), "biatan6", 90.0)], "spherical", 360.0, 192, 96))      # C3b shape
    add("ba6_sph_d3_rot", Job([FacetSpec(_cm(48, True), "biatan6", 90.0)], "spherical", 360.0, 160, 80, degree=3, **ROT))
    add("ba6_cube_d1", Job([FacetSpec(_cm(48, True), "biatan6", 90.0)], "cubemap", 90.0, 40, yaw=20.0))
    add("cm_sph_d1_support4_tile16", Job([FacetSpec(_cm(50), "cubemap", 90.0)], "spherical", 360.0, 128, 64,
                                         support_min=4, tile_size=16))
    # --- mounted single images of every projection ---------------------------------------
    add("rect_src_sph_d1", Job([FacetSpec(_rect(96, 64, 80.0, 30.0, 10.0, 5.0), "rectilinear", 80.0, yaw=30.0, pitch=10.0,
                                          roll=5.0)], "spherical", 360.0, 192, 96))
    add("rect_src_rect_d3", Job([FacetSpec(_rect(96, 64, 80.0), "rectilinear", 80.0)], "rectilinear", 60.0, 80, 60,
                                degree=3, yaw=8.0, roll=-12.0))
    add("fish_src_rect_d1", Job([FacetSpec(_ll(128, 128), "fisheye", 180.0, yaw=-20.0)], "rectilinear", 100.0, 80, 60))
    add("fish360_src_sph_d1", Job([FacetSpec(_ll(96, 96), "fisheye", 360.0)], "spherical", 360.0, 128, 64))
    add("ster_src_sph_d2", Job([FacetSpec(_ll(96, 64), "stereographic", 150.0, pitch=15.0)], "spherical", 360.0, 128, 64,
                               degree=2))
    add("cyl_src_fish_d1", Job([FacetSpec(_ll(128, 48), "cylindrical", 180.0)], "fisheye", 200.0, 72, 72, yaw=15.0))
    # --- channel counts -------------------------------------------------------------------
    add("grey_ll_rect_d3", Job([FacetSpec(_grey(_ll(128)), "spherical", 360.0)], "rectilinear", 90.0, 64, 36, degree=3,
                               **ROT))
    add("grey_cm_sph_d1_tw2", Job([FacetSpec(_grey(_cm(32)), "cubemap", 90.0)], "spherical", 360.0, 96, 48, twine=2))
    # --- multi-facet synopses -------------------------------------------------------------
    add("voronoi4_sph_d1", Job(_voronoi_facets(), "spherical", 360.0, 256, 128))                    # C5-B shape
    add("voronoi4_sph_d3_rot", Job(_voronoi_facets(), "spherical", 360.0, 192, 96, degree=3, **ROT))
    add("voronoi3_rect_d1_tw2", Job(_voronoi_facets(n=3), "rectilinear", 120.0, 96, 64, twine=2, yaw=-60.0))
    add("voronoi_mixed_sph_d1", Job([_ll_facet(128), FacetSpec(_rect(96, 64, 60.0, 40.0), "rectilinear", 60.0, yaw=40.0),
                                     FacetSpec(_cm(32), "cubemap", 90.0)], "spherical", 360.0, 160, 80))
    add("voronoi4_solo2", Job(_voronoi_facets(), "spherical", 360.0, 128, 64, solo=2))
    add("hdr3_rect_d1", Job(_bracket_facets(), "rectilinear", 70.0, 96, 64, yaw=20.0, synopsis="hdr_merge"))  # C5-A shape
    add("hdr3_sph_d3_tw2", Job(_bracket_facets(), "spherical", 120.0, 96, 48, yaw=20.0, degree=3, twine=2,
                               synopsis="hdr_merge"))
    add("lens3_voronoi_sph_d1", Job(_lens_facets(), "spherical", 360.0, 256, 128))
    add("lens1_rect_d3_tw2", Job(_lens_facets()[:1], "rectilinear", 60.0, 80, 60, yaw=-30.0, degree=3, twine=2))
    add("eev_voronoi_sph_d1", Job([FacetSpec(_rect(96, 64, 80.0, 0.0, 0.0, 0.0, 0.5), "rectilinear", 80.0, eev=13.0),
                                   FacetSpec(_rect(96, 64, 80.0, 60.0), "rectilinear", 80.0, yaw=60.0, eev=12.0),
                                   FacetSpec(_rect(96, 64, 80.0, -60.0), "rectilinear", 80.0, yaw=-60.0)],
                                  "spherical", 360.0, 192, 96))
    # --- alpha path: RGBA / grey+alpha facets, voronoi_syn_plus compositing, channel adaptation ---
    vf = _voronoi_facets(hfov=95.0, step=55.0)
    rgba = [FacetSpec(_rgba(f.image, k), f.projection, f.hfov, yaw=f.yaw, pitch=f.pitch, roll=f.roll)
            for k, f in enumerate(vf)]
    add("rgba1_rect_d1", Job(rgba[:1], "rectilinear", 90.0, 96, 64, yaw=-100.0))
    add("rgba1_sph_d3_tw2", Job(rgba[1:2], "spherical", 200.0, 128, 64, yaw=-45.0, degree=3, twine=2))
    add("rgba4_voronoi_sph_d1", Job(rgba, "spherical", 360.0, 256, 128))
    add("rgba4_voronoi_sph_d3", Job(rgba, "spherical", 360.0, 200, 100, degree=3, yaw=20.0, roll=5.0))
    add("rgba4_voronoi_rect_d1_tw2", Job(rgba, "rectilinear", 120.0, 96, 64, twine=2, yaw=-20.0))
    add("mixed_rgb_rgba_voronoi_sph_d1", Job([vf[0], rgba[1], vf[2], rgba[3]], "spherical", 360.0, 256, 128))
    add("mixed_grey_rgba_voronoi_sph_d1", Job([FacetSpec(_grey(vf[0].image), "rectilinear", 95.0, yaw=vf[0].yaw),
                                               rgba[1], FacetSpec(_ga(vf[2].image, 2), "rectilinear", 95.0,
                                                                  yaw=vf[2].yaw, pitch=vf[2].pitch, roll=vf[2].roll)],
                                              "spherical", 360.0, 256, 128))
    add("ga2_voronoi_sph_d1", Job([FacetSpec(_ga(f.image, k), f.projection, f.hfov, yaw=f.yaw, pitch=f.pitch,
                                             roll=f.roll) for k, f in enumerate(vf[:2])], "spherical", 360.0, 192, 96))
    add("ga_cm_sph_d3", Job([FacetSpec(_ga(_cm(32), 1), "cubemap", 90.0)], "spherical", 360.0, 128, 64, degree=3))
    add("ga1_rect_d5_tw2", Job([FacetSpec(_ga(vf[0].image, 0), "rectilinear", 95.0, yaw=vf[0].yaw)], "rectilinear", 90.0,
                               64, 48, yaw=vf[0].yaw, degree=5, twine=2))
    add("rgba_cm_sph_d1", Job([FacetSpec(_rgba(_cm(32), 1), "cubemap", 90.0)], "spherical", 360.0, 128, 64))
    hb = _bracket_facets()
    add("hdr3_rgba_rect_d1", Job([FacetSpec(_rgba(f.image, 1), f.projection, f.hfov, yaw=f.yaw, eev=f.eev) for f in hb],
                                 "rectilinear", 70.0, 96, 64, yaw=20.0, synopsis="hdr_merge"))
    # --- --mask_for: one facet painted white, the others black (masking.h); --nchannels reduces through mono_t ---
    add("maskfor1_voronoi4_sph_d1", Job(_voronoi_facets(), "spherical", 360.0, 256, 128, mask_for=1))
    add("maskfor2_voronoi4_grey_d3_tw2", Job(_voronoi_facets(), "spherical", 360.0, 128, 64, degree=3, twine=2, mask_for=2,
                                             out_channels=1))
    add("maskfor0_rgba4_voronoi_sph_d1", Job(rgba, "spherical", 360.0, 256, 128, mask_for=0))
    add("maskfor3_rgba4_ga_rect_d1", Job(rgba, "rectilinear", 120.0, 96, 64, yaw=-20.0, mask_for=3, out_channels=2))
    add("maskfor1_mixed_rgb_rgba_d1", Job([vf[0], rgba[1], vf[2], rgba[3]], "spherical", 360.0, 192, 96, mask_for=1,
                                          out_channels=2))
    add("maskfor0_cm_sph_d1", Job([FacetSpec(_cm(32), "cubemap", 90.0)], "spherical", 360.0, 96, 48, mask_for=0))
    add("maskfor1_hdr3_rect_d1", Job(_bracket_facets(), "rectilinear", 70.0, 96, 64, yaw=20.0, synopsis="hdr_merge",
                                     mask_for=1))
    # --- --nchannels on ordinary jobs: repix_t in both directions (environment.h:1205-1309) ---
    add("nch1_voronoi4_sph_d1", Job(_voronoi_facets(), "spherical", 360.0, 192, 96, out_channels=1))
    add("nch4_grey_ll_rect_d3", Job([FacetSpec(_grey(_ll(128)), "spherical", 360.0)], "rectilinear", 90.0, 64, 36, degree=3,
                                    out_channels=4, **ROT))
    add("nch3_rgba1_rect_d1_tw2", Job(rgba[:1], "rectilinear", 90.0, 96, 64, yaw=-100.0, twine=2, out_channels=3))
    add("nch2_hdr3_rect_d1", Job(_bracket_facets(), "rectilinear", 70.0, 96, 64, yaw=20.0, synopsis="hdr_merge", out_channels=2))
    # --- automatic twining (--twine omitted = -1): arguments::twine_setup picks the filter from the
    # magnification (envutil_main.cc:1450-1547) -------------------------------------------------
    add("auto_tw_down_ll_rect_d1", Job([_ll_facet(256)], "rectilinear", 100.0, 48, 32, twine=-1))          # mag < 1
    add("auto_tw_up_ll_rect_d1", Job([_ll_facet(128)], "rectilinear", 30.0, 96, 64, twine=-1, yaw=10.0))   # mag > 1, bilinear
    add("auto_tw_up_ll_rect_d3", Job([_ll_facet(128)], "rectilinear", 30.0, 96, 64, twine=-1, degree=3))   # mag > 1, cubic
    add("auto_tw_voronoi_d3", Job(_voronoi_facets(n=3), "spherical", 360.0, 512, 256, twine=-1, degree=3))  # several facets
    add("auto_tw_density_ll_ba6", Job([_ll_facet(256)], "biatan6", 90.0, 24, twine=-1, twine_density=1.5))
    # --- PTO exclude masks (k-lines) and lens crop (i-line S): alpha by polygon fill / crop + feather ---
    mf = _voronoi_facets(hfov=95.0, step=55.0)
    tri = ((10.5, 8.0), (60.0, 12.25), (30.0, 50.0))
    quad = ((70.0, 5.0), (90.0, 5.0), (90.0, 60.0), (70.0, 60.0))
    star = ((48.0, 2.0), (56.0, 40.0), (94.0, 30.0), (60.0, 50.0), (80.0, 63.0), (48.0, 52.0), (10.0, 62.0), (36.0, 46.0),
            (2.0, 20.0), (40.0, 36.0))
    add("mask1_rect_d1", Job([FacetSpec(mf[0].image, "rectilinear", 95.0, yaw=mf[0].yaw, masks=(tri, quad))],
                             "rectilinear", 90.0, 96, 64, yaw=mf[0].yaw))
    add("crop1_rect_d3", Job([FacetSpec(mf[1].image, "rectilinear", 95.0, yaw=mf[1].yaw, pitch=mf[1].pitch, roll=mf[1].roll,
                                        crop=(8, 80, 6, 50))], "rectilinear", 100.0, 96, 64, yaw=mf[1].yaw, degree=3))
    add("crop_fish_sph_d1", Job([FacetSpec(_ll(96, 96), "fisheye", 180.0, crop=(6, 92, 10, 84))], "spherical", 360.0,
                                192, 96))
    add("mask_crop4_voronoi_sph_d1", Job([FacetSpec(mf[0].image, "rectilinear", 95.0, yaw=mf[0].yaw, masks=(star,)),
                                          FacetSpec(mf[1].image, "rectilinear", 95.0, yaw=mf[1].yaw, pitch=mf[1].pitch,
                                                    roll=mf[1].roll, crop=(10, 90, 4, 60)),
                                          mf[2],
                                          FacetSpec(_rgba(mf[3].image, 3), "rectilinear", 95.0, yaw=mf[3].yaw,
                                                    pitch=mf[3].pitch, roll=mf[3].roll, masks=(tri,), crop=(0, 80, 0, 64))],
                                         "spherical", 360.0, 256, 128))
    add("mask_grey_sph_d1_tw2", Job([FacetSpec(_grey(mf[0].image), "rectilinear", 95.0, yaw=mf[0].yaw, masks=(quad,))],
                                    "spherical", 200.0, 128, 64, yaw=mf[0].yaw, twine=2))
    # --- 'W' windows: the file holds a crop of a larger image whose geometry the i-line describes ---
    big = _rect(128, 96, 90.0, 15.0, -5.0, 0.0)
    add("win_rect_sph_d1", Job([FacetSpec(np.ascontiguousarray(big[20:84, 16:112]), "rectilinear", 90.0, yaw=15.0, pitch=-5.0,
                                          window=(16, 112, 20, 84), total_width=128, total_height=96)],
                               "spherical", 360.0, 256, 128))
    add("win_rect_rect_d3_tw2", Job([FacetSpec(np.ascontiguousarray(big[8:72, 0:96]), "rectilinear", 90.0, yaw=15.0, pitch=-5.0,
                                               window=(0, 96, 8, 72), total_width=128, total_height=96)],
                                    "rectilinear", 80.0, 96, 64, yaw=15.0, degree=3, twine=2))
    add("win_voronoi_sph_d1", Job([FacetSpec(np.ascontiguousarray(big[20:84, 16:112]), "rectilinear", 90.0, yaw=15.0, pitch=-5.0,
                                             window=(16, 112, 20, 84), total_width=128, total_height=96),
                                   _voronoi_facets()[3]], "spherical", 360.0, 256, 128))
    # --- --single K: the target takes facet K's geometry, the result is un-brightened (C5 stage A) ---
    add("single0_hdr3_d1", Job(_bracket_facets(), "rectilinear", 70.0, 96, 64, synopsis="hdr_merge", single=0))
    add("single1_hdr3_d1", Job(_bracket_facets(), "rectilinear", 70.0, 96, 64, synopsis="hdr_merge", single=1))
    add("single2_voronoi4_d3_tw2", Job(_voronoi_facets(), "spherical", 360.0, 64, 32, single=2, degree=3, twine=2))
    add("single0_cm_ll", Job([FacetSpec(_cm(32), "cubemap", 90.0), _ll_facet(128)], "spherical", 360.0, 64, 32, single=0))
    tr = _translated_facets()
    # --single on facets WITH lens correction / shift / shear / translation: the generic stepper runs the
    # inverse planar transformation (inverse_lcp's spline) and the inverse translation (tf_ex_facet)
    lf, tf = _lens_facets(), _translated_facets()
    add("single0_lens3_d1", Job(lf, "spherical", 360.0, 64, 32, single=0))                  # a, b, c, d, e, g, t
    add("single1_lens3_d3_tw2", Job(lf, "spherical", 360.0, 64, 32, single=1, degree=3, twine=2))
    add("single2_lens3_d1", Job(lf, "spherical", 360.0, 64, 32, single=2))                  # fisheye target facet
    add("single0_tr3_d1", Job(tf, "spherical", 360.0, 64, 32, single=0))                    # target translation only
    add("single1_tr3_d1_tw2", Job(tf, "spherical", 360.0, 64, 32, single=1, twine=2))       # tilted translation plane
    add("single0_lens_tr_hdr_d1", Job([lf[0], tf[0], tf[1]], "spherical", 360.0, 64, 32, single=0, synopsis="hdr_merge"))
    add("single1_tr_lens_d1", Job([lf[1], tf[1], lf[2]], "spherical", 360.0, 64, 32, single=1))   # both sides translated
    # cropped output: PTO p-line with an S clause (the steppers see offset discrete coordinates)
    add("cropout_ll_sph_d1", Job([_ll_facet(256)], "spherical", 360.0, 1200, 600, crop_out=(500, 1120, 100, 420)))  # > 1 segment
    add("cropout_ll_rect_d3_tw2", Job([_ll_facet(128)], "rectilinear", 90.0, 160, 120, degree=3, twine=2,
                                      crop_out=(37, 150, 11, 97)))
    add("cropout_ll_cyl_d1_tw2", Job([_ll_facet(256)], "cylindrical", 360.0, 640, 90, twine=2, crop_out=(30, 610, 5, 70)))
    add("cropout_voronoi4_fish_d1", Job(_voronoi_facets(), "fisheye", 200.0, 200, 200, crop_out=(20, 180, 40, 200)))
    add("tr1_sph_d1", Job(tr[:1], "spherical", 360.0, 192, 96))
    add("tr1_rect_d1_tw2", Job(tr[1:2], "rectilinear", 100.0, 96, 64, yaw=30.0, twine=2))
    add("tr3_voronoi_sph_d1", Job(tr, "spherical", 360.0, 256, 128, yaw=11.0, pitch=3.0))
    add("tr3_voronoi_fish_d3_tw2", Job(tr, "fisheye", 200.0, 96, 96, degree=3, twine=2))
    add("tr2_cyl_d1", Job(tr[:2], "cylindrical", 300.0, 200, 60))
    add("tr2_ster_d1", Job(tr[1::-1], "stereographic", 220.0, 100, 100, yaw=-10.0))
    # translated facets with cubemap / biatan6 targets: the generic stepper over ir_to_ray_t / ba6_to_ray_t
    # (geometry.h:660-990; envutil_payload.cc:2097-2111 with STP = cubemap)
    add("tr1_cube_d1", Job(tr[:1], "cubemap", 90.0, 48, yaw=-20.0))
    add("tr3_voronoi_ba6_d1_tw2", Job(tr, "biatan6", 90.0, 40, twine=2, yaw=15.0, pitch=-8.0))
    add("tr2_cube100_d3", Job(tr[:2], "cubemap", 100.0, 36, degree=3))
    # `--single` on a wide translated facet: the target-side translation sends part of the rays behind the
    # translation plane, the generic stepper marks them (0, 0, -inf) and normalising makes that (0, 0, NaN). On x86 that
    # NaN carries a set sign bit, atan2 (0, NaN) is pi and every mounted image is missed (found by the random
    # sweep, seed 12 job 136: the kernels must not see the GPU's positive NaN as a hit at the image centre)
    rng = np.random.default_rng(136)
    nf = [FacetSpec(rng.random((20, 62, 3), dtype=np.float32), "spherical", 126.32447081741843, yaw=113.91373814521751),
          FacetSpec(rng.random((28, 56, 3), dtype=np.float32), "spherical", 360.0, yaw=-111.00090771641756,
                    pitch=70.69440350613604, roll=-10.717262957604202),
          FacetSpec(rng.random((21, 66, 3), dtype=np.float32), "spherical", 193.96977974745334, yaw=163.2839892925718),
          FacetSpec(rng.random((39, 58, 3), dtype=np.float32), "rectilinear", 113.6686449099416, yaw=-98.0931541828372,
                    pitch=-37.33135403223616, roll=-4.234541700616683, a=-0.0041605062769297756, b=0.014552337415172575,
                    c=-0.01240534595827941, tr_x=0.07629050379701123, tr_y=-0.03416623475390121, tr_z=0.036574020288230494)]
    add("single3_nan_rays_d0_tw2", Job(nf, "spherical", 360.0, 63, 17, degree=0, twine=2, twine_width=1.369265271160024,
                                       single=3))
    add("single3_nan_rays_d1", Job(nf, "spherical", 360.0, 63, 17, single=3))
    return J


JOBS = _build()
# the subset whose reference outputs are committed under tests/golden/
GOLDEN_JOBS = sorted(JOBS)


# Jobs that need command-line features outside the Job description: name -> (base job, contents of
# a .twf twining-filter file (x y weight per line, envutil_main.cc:1360-1403), extra arguments)
CLI_EXTRAS = {
    "twf_ll_rect_d1": ("ll_rect_d1_rot", "-0.3 -0.2 1\n0.3 -0.2 2\n0.0 0.35 3\n0.1 0.0 1.5\n", ["--twine_normalize"]),
    "twf_raw_voronoi_d1": ("voronoi4_sph_d1", "-0.25 0 0.5\n0.25 0 0.5\n", ["--twine_width", "1.5"]),
}


def _edge_jobs():
    """Degenerate sizes: one-pixel and ragged targets, sources smaller than a spline window, tiny cube
    faces. CPU: oracle == reference (golden). GPU: tests/test_gpu_parity.py::test_edge_jobs."""
    rng = np.random.default_rng(5)
    ll = lambda w, h: rng.random((h, w, 3), dtype=np.float32)
    cm = lambda f: rng.random((6 * f, f, 3), dtype=np.float32)
    E = {
        "edge_1x1_rect_from_ll8": Job([FacetSpec(ll(8, 4), "spherical", 360.0)], "rectilinear", 60.0, 1, 1),
        "edge_17x3_sph_from_cm4_d3": Job([FacetSpec(cm(4), "cubemap", 90.0)], "spherical", 360.0, 17, 3, degree=3),
        "edge_ll4x2_d3_rect": Job([FacetSpec(ll(4, 2), "spherical", 360.0)], "rectilinear", 90.0, 33, 9, degree=3),
        "edge_ll2x2_d3_sph": Job([FacetSpec(ll(2, 2), "spherical", 360.0)], "spherical", 360.0, 16, 8, degree=3),
        "edge_rect3x2_src_d2_tw2": Job([FacetSpec(rng.random((2, 3, 3), dtype=np.float32), "rectilinear", 50.0)],
                                       "rectilinear", 70.0, 31, 7, degree=2, twine=2),
        "edge_cube_target_w1": Job([FacetSpec(ll(16, 8), "spherical", 360.0)], "cubemap", 90.0, 1),
        # single-row / single-column rasters: zimt gates that axis as CONSTANT (always coordinate 0)
        "edge_ll2x1_src_d1": Job([FacetSpec(ll(2, 1), "spherical", 360.0)], "spherical", 360.0, 16, 8),
        "edge_rect1x5_src_d3": Job([FacetSpec(rng.random((5, 1, 3), dtype=np.float32), "rectilinear", 30.0)], "spherical",
                                   360.0, 16, 8, degree=3),
        "edge_513_wide_cyl_tw3": Job([FacetSpec(ll(32, 16), "spherical", 360.0)], "cylindrical", 360.0, 513, 2, twine=3),
    }
    for k, v in E.items():
        v.name = k
    return E


EDGE_JOBS = _edge_jobs()


def _odd_cube_jobs():
    """Cubemaps with an ODD face width: the support fill then reads frame texels it is rewriting, so its result
    depends on the order zimt::process works in (16-pixel vectors, line after line). The
    oracle restates that order and equals the reference wherever the reference is deterministic; the kernels
    (stage.cu, k_cm_fill_ordered) are held to the oracle."""
    rng = np.random.default_rng(9)
    cm = lambda f: rng.random((6 * f, f, 3), dtype=np.float32)
    O = {
        "odd_cm5_sph_d1": Job([FacetSpec(cm(5), "cubemap", 90.0)], "spherical", 360.0, 64, 32),
        "odd_cm35_sph_d1": Job([FacetSpec(cm(35), "cubemap", 90.0)], "spherical", 360.0, 160, 80),
        "odd_cm47_rect_d3": Job([FacetSpec(cm(47), "cubemap", 90.0)], "rectilinear", 100.0, 96, 64, degree=3, yaw=40.0,
                                pitch=50.0),
        "odd_ba9_sph_d1_tw2": Job([FacetSpec(cm(9), "biatan6", 90.0)], "spherical", 360.0, 48, 24, twine=2),
        "odd_cm101_hfov96_sph_d1": Job([FacetSpec(cm(101), "cubemap", 96.0)], "spherical", 360.0, 200, 100),
    }
    for k, v in O.items():
        v.name = k
    return O


ODD_CUBE_JOBS = _odd_cube_jobs()

# --split FORMAT runs of the reference CLI (one output per facet, the solo facet excepted): name ->
# (base job, extra command-line arguments). Golden: manifest[name]["outputs"][facet] = shape + sha256.
SPLITS = {
    "split_lens3_d1": ("lens3_voronoi_sph_d1", []),
    "split_tr3_solo2_d1_tw2": ("tr3_voronoi_sph_d1", ["--solo", "2", "--twine", "2"]),
}

# tethered output (to_screen_t, envutil_payload.cc:298-413): jobs whose uint32 sRGBA frame is pinned by the reference's
# own tethered pipeline (tests/golden/screen.json, tools/make_golden_screen.py). One to four channels, twining, both
# synopses, and a 'single' job with a gain (which the tethered path does not apply, envutil_payload.cc:491).
SCREEN_JOBS = ["ll_rect_d1", "cm_sph_d3", "ll_fish_d1_tw4", "voronoi4_sph_d1", "hdr3_rect_d1", "grey_ll_rect_d3",
               "ga_cm_sph_d3", "ga2_voronoi_sph_d1", "rgba1_rect_d1", "rgba4_voronoi_sph_d1", "single1_hdr3_d1"]
