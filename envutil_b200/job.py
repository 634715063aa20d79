"""Job description shared by the Python host API, the tests and the bench.

A Job is what envutil's command line describes (reference envutil_main.cc:190-372): a set of
facets (source images with projection, field of view and orientation), a target projection and
size, and the interpolation / twining / synopsis options. `cli_args()` spells the job for the
reference binary, `structs()` marshals it into the POD structs of include/envutil_b200.h using
the library's own host-side set-up functions.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi


@dataclass
class FacetSpec:
    image: Optional[np.ndarray]  # H x W x C float32 (may be None when only geometry is needed)
    projection: str
    hfov: float  # degrees
    yaw: float = 0.0
    pitch: float = 0.0
    roll: float = 0.0
    brighten: float = 1.0  # linear gain; overridden by eev when the job goes through PTO
    eev: float = 0.0       # PTO Eev (0 = not given); brighten = 2^(Eev - mean Eev), envutil_main.cc:1006-1061
    a: float = 0.0
    b: float = 0.0
    c: float = 0.0
    d: float = 0.0  # PTO 'd' (horizontal shift, pixels)
    e: float = 0.0  # PTO 'e' (vertical shift, pixels)
    g: float = 0.0  # shear
    t: float = 0.0
    tr_x: float = 0.0
    tr_y: float = 0.0
    tr_z: float = 0.0
    tp_y: float = 0.0
    tp_p: float = 0.0
    window: Optional[tuple] = None  # PTO i-line W clause (x0, x1, y0, y1): `image` is this window of a larger
                                    # image whose total size is (total_width, total_height)
    total_width: int = 0
    total_height: int = 0
    crop: Optional[tuple] = None   # PTO i-line S clause: (x0, x1, y0, y1) lens crop -> alpha (needs the PTO route)
    masks: tuple = ()              # PTO k-lines, variant t0: polygons ((x, y), ...) excluded -> alpha
    width: int = 0   # used when image is None
    height: int = 0
    nchannels: int = 3

    def native_shape(self):
        if self.image is not None:
            h, w = self.image.shape[:2]
            c = 1 if self.image.ndim == 2 else self.image.shape[2]
            return w, h, c
        return self.width, self.height, self.nchannels

    def has_alpha_spec(self):
        return self.crop is not None or len(self.masks) > 0

    def shape(self):
        """Shape of the STAGED facet: masks / lens crop add an alpha channel to 1- and 3-channel
        images (envutil_main.cc:1065-1069)."""
        w, h, c = self.native_shape()
        if self.has_alpha_spec() and c in (1, 3):
            c += 1
        return w, h, c

    def alpha_spec(self):
        """ctypes AlphaSpec for the C ABI / the oracle (keeps the arrays alive on the object)."""
        a = capi.AlphaSpec()
        a.native_nchannels = self.native_shape()[2]
        if self.crop is not None:
            a.has_crop = 1
            a.crop_x0, a.crop_x1, a.crop_y0, a.crop_y1 = (int(v) for v in self.crop)
        a.n_masks = len(self.masks)
        sizes = (C.c_int32 * max(1, len(self.masks)))(*[len(m) for m in self.masks])
        flat = [float(np.float32(v)) for m in self.masks for xy in m for v in xy]
        xy = (C.c_float * max(1, len(flat)))(*flat)
        a.mask_sizes, a.mask_xy = sizes, xy
        a._keep = (sizes, xy)
        return a


@dataclass
class Job:
    facets: List[FacetSpec]
    projection: str
    hfov: float  # degrees
    width: int
    height: int = 0
    yaw: float = 0.0
    pitch: float = 0.0
    roll: float = 0.0
    degree: int = 1
    prefilter: int = -1
    twine: int = 0
    twine_width: float = 1.0
    twine_density: float = 1.0
    twine_sigma: float = 0.0
    twine_threshold: float = 0.0
    twine_max: int = 8
    synopsis: str = "panorama"
    solo: int = -1
    mask_for: int = -1      # --mask_for K: facet K is painted white, all others black (envutil_main.cc:999-1001,1077-1091)
    out_channels: int = 0   # --nchannels: the job's channel count instead of the maximum over the facets (:1131-1133)
    crop_out: Optional[tuple] = None  # PTO p-line S clause (x0, x1, y0, y1): only this window of the target is
                                      # rendered and stored (envutil_main.cc:615-627, envutil_payload.cc:440-474)
    single: int = -1  # --single K: render into facet K's geometry, undo its brighten (envutil_main.cc:1157-1178)
    support_min: int = 8
    tile_size: int = 64
    padded: Optional[bool] = None  # back-end option, eu_opts.reserved[0]: 16-byte RGB texels in HBM - None = the library's
                                   # rule (bilinear and nearest-neighbour jobs), True = always, False = never
    no_tiles: bool = False  # back-end option: never stage gather footprints in shared memory (reserved[1] bit 0)
    narrow_stores: bool = False  # back-end option: 4-byte pixel stores even into peer frames (reserved[1] bit 2)
    no_spec: bool = False   # back-end option: never use the kernels compiled for one job shape (reserved[1] bit 1)
    contracted: Optional[bool] = None  # arithmetic of the render kernels: fused multiply-adds in the window evaluation
                                       # (EU_OPT_CONTRACTED); None = what EU_ARITHMETIC says (default: exact)
    name: str = ""

    # ---- reference command line (real spellings, envutil_main.cc:190-372) ----
    def uses_pto(self):
        return self.crop_out is not None or any(
            f.eev or f.a or f.b or f.c or f.d or f.e or f.g or f.t or f.tr_x or f.tr_y or f.tr_z
            or f.has_alpha_spec() or f.window is not None for f in self.facets)

    _PTO_CODE = {"rectilinear": 0, "cylindrical": 1, "fisheye": 3, "spherical": 4, "stereographic": 10}
    _PTO_P_CODE = {"rectilinear": 0, "cylindrical": 1, "spherical": 2, "fisheye": 3, "stereographic": 4}  # p-line, :585-602

    def pto_lines(self, facet_paths):
        """i-lines for the facets (PTO subset, reference envutil_main.cc:655-822)."""
        lines = []
        for f, p in zip(self.facets, facet_paths):
            w, h, _ = f.native_shape()
            if f.window is not None:
                w, h = f.total_width, f.total_height
            ln = (f'i w{w} h{h} f{self._PTO_CODE[f.projection]} v{float(f.hfov)!r} y{float(f.yaw)!r} '
                  f'p{float(f.pitch)!r} r{float(f.roll)!r}')
            for key, val in (("Eev", f.eev), ("a", f.a), ("b", f.b), ("c", f.c), ("d", f.d), ("e", f.e), ("g", f.g),
                             ("t", f.t), ("TrX", f.tr_x), ("TrY", f.tr_y), ("TrZ", f.tr_z), ("Tpy", f.tp_y),
                             ("Tpp", f.tp_p)):
                if val:
                    ln += f" {key}{float(val)!r}"
            if f.crop is not None:
                ln += " S%d,%d,%d,%d" % tuple(int(v) for v in f.crop)
            if f.window is not None:
                ln += " W%d,%d,%d,%d" % tuple(int(v) for v in f.window)
            ln += f' n"{p}"'
            lines.append(ln)
        for i, f in enumerate(self.facets):
            for m in f.masks:  # k-lines: exclude masks (envutil_main.cc:829-904)
                pts = " ".join("%s %s" % (repr(float(x)), repr(float(y))) for x, y in m)
                lines.append(f'k i{i} t0 p"{pts}"')
        if self.crop_out is not None:
            # the p-line is honoured only when --width is absent; then it supplies projection, size and
            # hfov, and the camera angles are NOT converted from degrees (envutil_main.cc:1180-1194)
            assert self.yaw == 0 and self.pitch == 0 and self.roll == 0 and self.height and self.single < 0
            lines.append("p f%d w%d h%d v%r S%d,%d,%d,%d" % ((self._PTO_P_CODE[self.projection], self.width, self.height,
                                                            float(self.hfov)) + tuple(int(v) for v in self.crop_out)))
        return lines

    def cli_args(self, facet_paths, output):
        args = []
        if self.uses_pto():
            for ln in self.pto_lines(facet_paths):
                args += ["--pto_line", ln]
        else:
            for f, p in zip(self.facets, facet_paths):
                args += ["--facet", p, f.projection, repr(float(f.hfov)), repr(float(f.yaw)), repr(float(f.pitch)),
                         repr(float(f.roll))]
        if self.crop_out is None:
            args += ["--projection", self.projection, "--hfov", repr(float(self.hfov)), "--width", str(self.width)]
            if self.height:
                args += ["--height", str(self.height)]
        args += ["--yaw", repr(float(self.yaw)), "--pitch", repr(float(self.pitch)), "--roll", repr(float(self.roll))]
        args += ["--degree", str(self.degree), "--prefilter", str(self.prefilter), "--twine", str(self.twine)]
        if self.twine != 0:
            args += ["--twine_width", repr(float(self.twine_width)), "--twine_sigma", repr(float(self.twine_sigma)),
                     "--twine_threshold", repr(float(self.twine_threshold))]
            if self.twine_density != 1.0:
                args += ["--twine_density", repr(float(self.twine_density))]
            if self.twine_max != 8:
                args += ["--twine_max", str(self.twine_max)]
        if self.synopsis != "panorama":
            args += ["--synopsis", self.synopsis]
        if self.solo >= 0:
            args += ["--solo", str(self.solo)]
        if self.mask_for >= 0:
            args += ["--mask_for", str(self.mask_for)]
        if self.out_channels:
            args += ["--nchannels", str(self.out_channels)]
        if self.single >= 0:
            args += ["--single", str(self.single)]
        if self.support_min != 8:
            args += ["--support_min", str(self.support_min)]
        if self.tile_size != 64:
            args += ["--tile_size", str(self.tile_size)]
        args += ["--output", output]
        return args

    def facet_gains(self):
        """brighten per facet as arguments::init derives it (envutil_main.cc:1006-1061): from the
        PTO Eev values when any is given (facet_spec::brighten is a float), else 1."""
        eevs = [np.float32(f.eev) for f in self.facets]
        given = [e for e in eevs if e != 0]
        if not given or not self.uses_pto():
            return [float(np.float32(f.brighten)) for f in self.facets]
        acc = np.float32(0.0)  # `float eev_sum`, envutil_main.cc:519
        for e in given:
            acc = np.float32(acc + e)
        mean = np.float32(acc / np.float32(len(given)))
        return [1.0 if e == 0 else float(np.float32(2.0 ** float(np.float32(e - mean)))) for e in eevs]

    # ---- POD structs of the C ABI ----
    def structs(self, lib=None):
        lib = lib or capi.load()
        # channel-count rule of arguments::init (envutil_main.cc:1063-1122): the maximum over the
        # facets; RGB together with any alpha-carrying facet renders RGBA
        counts = [f.shape()[2] for f in self.facets]
        nch = max(counts)
        if nch == 3 and any(c in (2, 4) for c in counts):
            nch = 4
        if self.out_channels:
            nch = self.out_channels
        t = capi.Target()
        t.projection = capi.PROJECTION_NAMES.index(self.projection)
        t.width, t.height, t.nchannels = self.width, self.height, nch
        # the reference parses the target's --hfov/--yaw/--pitch/--roll as FLOAT (ap[...].get<float>,
        # envutil_main.cc:468-477) and converts to radians in double (:1199-1202)
        f32 = lambda v: float(np.float32(v)) * (math.pi / 180.0)
        t.hfov = f32(self.hfov)
        t.yaw, t.pitch, t.roll = f32(self.yaw), f32(self.pitch), f32(self.roll)
        if self.crop_out is not None:  # p-line route: hfov is parsed as a double there (envutil_main.cc:613)
            t.hfov = float(self.hfov) * (math.pi / 180.0)
            x0, x1, y0, y1 = (int(v) for v in self.crop_out)
            t.crop_x0, t.crop_y0, t.crop_width, t.crop_height = x0, y0, x1 - x0, y1 - y0
        if self.single < 0:
            capi.check(lib.eu_target_prepare(C.byref(t)), lib)
        n = len(self.facets)
        fa = (capi.Facet * n)()
        gains = self.facet_gains()
        for i, f in enumerate(self.facets):
            w, h, c = f.shape()
            s = fa[i]
            s.projection = capi.PROJECTION_NAMES.index(f.projection)
            s.width, s.height, s.nchannels = w, h, c
            if f.window is not None:  # 'W' clause: geometry from the total size, raster = the window
                x0, x1, y0, y1 = (int(v) for v in f.window)
                s.width, s.height = f.total_width, f.total_height
                s.window_x_offset, s.window_y_offset = x0, y0
                s.window_width, s.window_height = x1 - x0, y1 - y0
                assert (s.window_width, s.window_height) == (w, h), "image must have the window's size"
                w, h = s.width, s.height
            # facet angles are parsed as doubles (%F / std::stod) and scaled by M_PI / 180.0
            s.hfov = f.hfov * (math.pi / 180.0)
            s.yaw, s.pitch, s.roll = (v * (math.pi / 180.0) for v in (f.yaw, f.pitch, f.roll))
            s.tr_x, s.tr_y, s.tr_z = f.tr_x, f.tr_y, -f.tr_z  # TrZ is negated (envutil_main.cc:787-789)
            s.tp_y, s.tp_p = (math.pi / 180.0) * f.tp_y, (math.pi / 180.0) * f.tp_p
            s.shear_g, s.shear_t = f.g / h, f.t / w  # envutil_main.cc:795-796
            s.a, s.b, s.c, s.h, s.v = f.a, f.b, f.c, f.d, f.e
            s.brighten = gains[i]
            if self.mask_for >= 0:  # facet_spec::masked 1 / 0 -> eu_facet_t.masked 2 (white) / 1 (black)
                s.masked = 2 if i == self.mask_for else 1
            capi.check(lib.eu_facet_prepare(C.byref(s)), lib)
        if self.single >= 0:  # (facet_base&) args = facet_spec_v[single]: geometry in radians as the facet has it
            sf = fa[self.single]
            t.single = self.single + 1
            t.projection, t.width, t.height = sf.projection, sf.width, sf.height
            t.hfov, t.yaw, t.pitch, t.roll = sf.hfov, sf.yaw, sf.pitch, sf.roll
            b = np.float32(gains[self.single])
            t.gain = float(np.float32(1.0 / float(b))) if b != 1.0 else 0.0
            capi.check(lib.eu_target_prepare(C.byref(t)), lib)
        o = capi.Opts()
        o.spline_degree = self.degree
        o.prefilter_degree = self.prefilter
        o.synopsis = capi.SYN_HDR_MERGE if self.synopsis == "hdr_merge" else capi.SYN_PANORAMA
        o.solo = 0 if n == 1 else self.solo  # forced for a single facet (envutil_main.cc:996-997)
        o.support_min, o.tile_size = self.support_min, self.tile_size
        o.reserved[0] = 0 if self.padded is None else (1 if self.padded else 2)
        contracted = capi.ARITHMETIC == "contracted" if self.contracted is None else self.contracted
        o.reserved[1] = (capi.OPT_NO_TILES if self.no_tiles else 0) | (capi.OPT_NO_SHAPES if self.no_spec else 0) | \
            (capi.OPT_NARROW_STORES if self.narrow_stores else 0) | (capi.OPT_CONTRACTED if contracted else 0)
        taps = (capi.Tap * 1024)()
        tw = C.c_int(0)
        ntaps = lib.eu_make_spread(C.byref(t), C.byref(o), n, fa, self.twine, self.twine_width, self.twine_density,
                                   self.twine_sigma, self.twine_threshold, self.twine_max, taps, 1024, C.byref(tw))
        if ntaps < 0:
            capi.check(ntaps, lib)
        return t, fa, o, taps, ntaps
