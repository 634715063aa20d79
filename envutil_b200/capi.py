"""ctypes view of the C ABI declared in include/envutil_b200.h.

This is plumbing for the Python tests, the bench and the smoke test: it loads the in-tree
shared library built by __graft_entry__.build() and mirrors the POD structs field for field.
There is no fallback: if the library is missing, loading raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EU_ARITHMETIC=contracted makes Job.structs() ask for the contracted arithmetic (EU_OPT_CONTRACTED: fused
# multiply-adds in the window evaluation) unless a Job says otherwise; anything else is the bit-exact default.
# Both arithmetics live in the one library.
ARITHMETIC = "contracted" if os.environ.get("EU_ARITHMETIC", "") == "contracted" else "exact"
LIB_PATH = os.path.join(_HERE, "libenvutil_b200.so")
OPT_NO_TILES, OPT_NO_SHAPES, OPT_NARROW_STORES, OPT_CONTRACTED = 1, 2, 4, 16  # eu_opts_t.reserved[1]

# eu_projection_t (reference envutil_basic.h:99-109)
SPHERICAL, CYLINDRICAL, RECTILINEAR, STEREOGRAPHIC, FISHEYE, CUBEMAP, BIATAN6, PRJ_NONE = range(8)
PROJECTION_NAMES = ["spherical", "cylindrical", "rectilinear", "stereographic", "fisheye", "cubemap", "biatan6"]
SYN_PANORAMA, SYN_HDR_MERGE = 0, 1
EU_OK = 0


class Facet(C.Structure):
    _fields_ = [
        ("projection", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("nchannels", C.c_int32),
        ("hfov", C.c_double), ("yaw", C.c_double), ("pitch", C.c_double), ("roll", C.c_double),
        ("tr_x", C.c_double), ("tr_y", C.c_double), ("tr_z", C.c_double),
        ("tp_y", C.c_double), ("tp_p", C.c_double), ("tp_r", C.c_double),
        ("shear_g", C.c_double), ("shear_t", C.c_double),
        ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
        ("h", C.c_double), ("v", C.c_double), ("brighten", C.c_double),
        ("x0", C.c_double), ("x1", C.c_double), ("y0", C.c_double), ("y1", C.c_double),
        ("step", C.c_double),
        ("s", C.c_double), ("d", C.c_double), ("r_max", C.c_double), ("cap_radius", C.c_double),
        ("shift_h", C.c_double), ("shift_v", C.c_double),
        ("has_shift", C.c_int32), ("has_lcp", C.c_int32), ("has_shear", C.c_int32),
        ("has_2d_tf", C.c_int32), ("has_translation", C.c_int32),
        ("window_width", C.c_int32), ("window_height", C.c_int32),
        ("window_x_offset", C.c_int32), ("window_y_offset", C.c_int32),
        ("masked", C.c_int32),
    ]


class Target(C.Structure):
    _fields_ = [
        ("projection", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("nchannels", C.c_int32),
        ("hfov", C.c_double), ("yaw", C.c_double), ("pitch", C.c_double), ("roll", C.c_double),
        ("gain", C.c_double),
        ("x0", C.c_double), ("x1", C.c_double), ("y0", C.c_double), ("y1", C.c_double),
        ("step", C.c_double),
        ("crop_x0", C.c_int32), ("crop_y0", C.c_int32), ("crop_width", C.c_int32), ("crop_height", C.c_int32),
        ("single", C.c_int32), ("reserved", C.c_int32),
    ]

    def out_shape(self):
        """(rows, columns) of the raster a job produces: the target, or its crop."""
        if self.crop_width > 0:
            return self.crop_height, self.crop_width
        return self.height, self.width


class Opts(C.Structure):
    _fields_ = [
        ("spline_degree", C.c_int32), ("prefilter_degree", C.c_int32), ("synopsis", C.c_int32),
        ("solo", C.c_int32), ("support_min", C.c_int32), ("tile_size", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


class Tap(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("w", C.c_float)]


class Timing(C.Structure):
    _fields_ = [("render_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("launches", C.c_int32), ("shape", C.c_int32)]


class AlphaSpec(C.Structure):
    _fields_ = [("native_nchannels", C.c_int32), ("has_crop", C.c_int32), ("crop_x0", C.c_int32),
                ("crop_x1", C.c_int32), ("crop_y0", C.c_int32), ("crop_y1", C.c_int32), ("n_masks", C.c_int32),
                ("mask_sizes", C.POINTER(C.c_int32)), ("mask_xy", C.POINTER(C.c_float))]


SourceH = C.c_void_p
_FP = C.POINTER(C.c_float)

# every symbol include/envutil_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "eu_get_vfov": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_double]),
    "eu_get_step": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_double]),
    "eu_get_extent": (None, [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]),
    "eu_facet_prepare": (C.c_int, [C.POINTER(Facet)]),
    "eu_target_prepare": (C.c_int, [C.POINTER(Target)]),
    "eu_rotation_matrix": (None, [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "eu_facet_basis": (None, [C.POINTER(Target), C.POINTER(Facet), C.POINTER(C.c_double)]),
    "eu_make_spread": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet), C.c_int, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_int, C.POINTER(Tap), C.c_int, C.POINTER(C.c_int)]),
    "eu_cubemap_metrics": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_double)]),
    "eu_init": (C.c_int, [C.c_int]),
    "eu_shutdown": (None, []),
    "eu_last_error": (C.c_char_p, []),
    "eu_device_count": (C.c_int, []),
    "eu_render_arithmetic": (C.c_int, []),
    "eu_source_upload": (C.c_int, [C.c_char_p, C.POINTER(Facet), C.POINTER(Opts), C.c_void_p,
                                   C.POINTER(SourceH), C.POINTER(Timing)]),
    "eu_source_upload_device": (C.c_int, [C.c_char_p, C.POINTER(Facet), C.POINTER(Opts), C.c_void_p, C.c_void_p,
                                          C.POINTER(SourceH), C.POINTER(Timing)]),
    "eu_source_upload_alpha": (C.c_int, [C.c_char_p, C.POINTER(Facet), C.POINTER(Opts), C.c_void_p,
                                         C.POINTER(AlphaSpec), C.POINTER(SourceH), C.POINTER(Timing)]),
    "eu_source_find": (SourceH, [C.c_char_p]),
    "eu_source_release": (C.c_int, [SourceH]),
    "eu_cycle": (C.c_int, []),
    "eu_source_container_floats": (C.c_size_t, [SourceH, C.POINTER(C.c_int32)]),
    "eu_source_download": (C.c_int, [SourceH, C.c_void_p]),
    "eu_render": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet), C.POINTER(SourceH),
                            C.POINTER(Tap), C.c_int, C.c_void_p, C.POINTER(Timing)]),
    "eu_render_rows": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet),
                                 C.POINTER(SourceH), C.POINTER(Tap), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.POINTER(Timing)]),
    "eu_render_rows_pitched": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet),
                                         C.POINTER(SourceH), C.POINTER(Tap), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_int, C.c_void_p, C.POINTER(Timing)]),
    "eu_render_rect_pitched": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet),
                                         C.POINTER(SourceH), C.POINTER(Tap), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(Timing)]),
    "eu_source_reserve": (C.c_int, [C.c_char_p, C.POINTER(Facet), C.POINTER(Opts), C.POINTER(SourceH),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "eu_source_commit": (C.c_int, [SourceH, C.POINTER(Facet), C.POINTER(Opts), C.c_void_p, C.POINTER(Timing)]),
    "eu_source_write_rect": (C.c_int, [SourceH, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eu_source_upload_async": (C.c_int, [C.c_char_p, C.POINTER(Facet), C.POINTER(Opts), C.c_void_p,
                                         C.POINTER(SourceH)]),
    "eu_render_async": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet), C.POINTER(SourceH),
                                  C.POINTER(Tap), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "eu_job_wait": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "eu_frame_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "eu_frame_free": (C.c_int, [C.c_void_p]),
    "eu_frame_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "eu_frame_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "eu_frame_close": (C.c_int, [C.c_void_p]),
    "eu_screen_lut": (None, [C.POINTER(C.c_float)]),
    "eu_render_screen": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet), C.POINTER(SourceH),
                                   C.POINTER(Tap), C.c_int, C.c_void_p, C.POINTER(Timing)]),
    "eu_to_screen_device": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]),
    "eu_debug_planes": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet),
                                  C.POINTER(SourceH), C.c_void_p]),
    "eu_debug_tie_plane": (C.c_int, [C.POINTER(Target), C.POINTER(Opts), C.c_int, C.POINTER(Facet),
                                     C.POINTER(SourceH), C.c_int, C.c_void_p]),
}

_lib = None


def load():
    """Load libenvutil_b200.so (built in-tree). Raises if it is missing or a symbol is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.eu_render_arithmetic() != 2:
        raise RuntimeError(f"{LIB_PATH} does not carry both arithmetics: rebuild (__graft_entry__.build())")
    _lib = lib
    return lib


def check(rc, lib=None):
    if rc != EU_OK:
        lib = lib or load()
        raise RuntimeError(f"envutil_b200: status {rc}: {lib.eu_last_error().decode()}")
