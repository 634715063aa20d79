"""Test-side plumbing: ctypes binding of the C oracle (oracle/liboracle.so), a runner for the
unmodified reference binaries under oracle/_ref/ and comparison helpers.

Nothing here is product code. /root/reference is never read at run time: the reference
binaries are self-contained executables built by oracle/Makefile.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from envutil_b200 import capi, euf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_binary(kind="pm"):
    name = {"pm": "envutil_ref_pm", "libm": "envutil_ref", "fast": "envutil_ref_fast"}[kind]
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


_orc = None


def oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        lib = C.CDLL(ORACLE_SO)
        lib.orc_source_create.restype = C.c_void_p
        lib.orc_source_create.argtypes = [C.POINTER(capi.Facet), C.POINTER(capi.Opts), C.c_void_p]
        lib.orc_source_free.argtypes = [C.c_void_p]
        lib.orc_source_create_alpha.restype = C.c_void_p
        lib.orc_source_create_alpha.argtypes = [C.POINTER(capi.Facet), C.POINTER(capi.Opts), C.c_void_p,
                                                C.POINTER(capi.AlphaSpec)]
        lib.orc_source_container.restype = C.POINTER(C.c_float)
        lib.orc_source_container.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        lib.orc_render.restype = C.c_int
        lib.orc_render.argtypes = [C.POINTER(capi.Target), C.POINTER(capi.Opts), C.c_int, C.POINTER(capi.Facet),
                                   C.POINTER(C.c_void_p), C.POINTER(capi.Tap), C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int]
        lib.orc_set_arithmetic.argtypes = [C.c_int]
        lib.orc_get_extent.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
        lib.orc_get_step.restype = C.c_double
        lib.orc_get_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        lib.orc_rotation.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
        lib.orc_poles.restype = C.c_int
        lib.orc_poles.argtypes = [C.c_int, C.POINTER(C.c_longdouble)]
        lib.orc_weight_matrix.argtypes = [C.c_int, C.POINTER(C.c_float)]
        _orc = lib
    return _orc


def oracle_sources(job, structs=None):
    """Stage every facet of the job with the oracle; returns (handles array, list for freeing)."""
    lib = oracle()
    t, fa, o, taps, ntaps = structs or job.structs()
    hs = (C.c_void_p * len(job.facets))()
    for i, f in enumerate(job.facets):
        img = np.ascontiguousarray(f.image, dtype=np.float32)
        if f.has_alpha_spec():
            a = f.alpha_spec()
            hs[i] = lib.orc_source_create_alpha(C.byref(fa[i]), C.byref(o), img.ctypes.data, C.byref(a))
        else:
            hs[i] = lib.orc_source_create(C.byref(fa[i]), C.byref(o), img.ctypes.data)
    return hs


def oracle_container(handle):
    lib = oracle()
    shp = (C.c_int32 * 4)()
    p = lib.orc_source_container(handle, shp)
    return p, tuple(shp)


def oracle_render(job, want_index=False, threads=0, rows=None, sources=None, contracted=False):
    """Render the job with the C oracle. Returns H x W x C float32 (and the index plane).
    contracted: the arithmetic of the opt-in libenvutil_b200_fma.so build (orc_set_arithmetic) instead of
    the reference's."""
    lib = oracle()
    st = job.structs()
    t, fa, o, taps, ntaps = st
    hs = sources if sources is not None else oracle_sources(job, st)
    oh, ow = t.out_shape()
    row0, row1 = rows or (0, oh)
    out = np.empty((row1 - row0, ow, t.nchannels), dtype=np.float32)
    idx = np.empty((row1 - row0, ow), dtype=np.int32) if want_index else None
    lib.orc_set_arithmetic(1 if contracted else 0)
    try:
        rc = lib.orc_render(C.byref(t), C.byref(o), len(job.facets), fa, hs, taps, ntaps, row0, row1,
                            out.ctypes.data, idx.ctypes.data if want_index else None, threads)
    finally:
        lib.orc_set_arithmetic(0)
    assert rc == 0, rc
    if sources is None:
        for h in hs:
            lib.orc_source_free(h)
    return (out, idx) if want_index else out


def reference_render(job, kind="pm", extra_args=(), keep_log=False):
    """Run the job through the unmodified reference CLI (oracle/_ref). Returns H x W x C."""
    exe = ref_binary(kind)
    assert exe, "oracle/_ref is not built (make -C oracle ref)"
    with tempfile.TemporaryDirectory(prefix="euref_") as d:
        paths = []
        for i, f in enumerate(job.facets):
            p = os.path.join(d, f"facet{i}.euf")
            euf.write_euf(p, f.image)
            paths.append(p)
        outp = os.path.join(d, "out.euf")
        cmd = [exe] + job.cli_args(paths, outp) + list(extra_args)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(outp):
            raise RuntimeError(f"reference failed ({r.returncode}): {' '.join(cmd)}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        img = euf.read_euf(outp)
        return (img, r.stdout) if keep_log else img


def reference_screen(job, kind="pm", refine=False):
    """The job through the reference's TETHERED pipeline (handle_job -> core(tethered) -> act + to_screen_t,
    envutil_main.cc:1755-1868, envutil_payload.cc:298-413,524-531): H x W uint32 sRGBA, as visor's frame buffer
    receives it. The oracle build's visor stub (oracle/shim/visor_stub/visor.h) runs one job described by the
    environment instead of visor's shared-memory queue."""
    exe = ref_binary(kind)
    assert exe, "oracle/_ref is not built (make -C oracle ref)"
    t = job.structs()[0]
    oh, ow = t.out_shape()
    with tempfile.TemporaryDirectory(prefix="euref_") as d:
        paths = []
        for i, f in enumerate(job.facets):
            p = os.path.join(d, f"facet{i}.euf")
            euf.write_euf(p, f.image)
            paths.append(p)
        argf, outp = os.path.join(d, "args.txt"), os.path.join(d, "frame.u32")
        with open(argf, "w") as fh:
            fh.write("\n".join(["visor"] + job.cli_args(paths, os.path.join(d, "none.euf"))) + "\n")
        env = dict(os.environ, EU_TETHER_SPEC="%d %d %r %r %r %r 1.0 %d" % (ow, oh, float(job.yaw), float(job.pitch),
                                                                          float(job.roll), float(job.hfov), int(refine)),
                   EU_TETHER_ARGS=argf, EU_TETHER_OUT=outp)
        r = subprocess.run([exe, "+"], capture_output=True, text=True, env=env)
        if r.returncode != 0 or not os.path.exists(outp):
            raise RuntimeError(f"reference (tethered) failed ({r.returncode})\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        return np.fromfile(outp, dtype="<u4").reshape(oh, ow)


def compare(a, b, eps=1e-3):
    """max and RMS of |a-b| / max(|b|, eps), plus the count of differing values."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), eps)
    return {"max_rel": float(rel.max()), "rms_rel": float(np.sqrt((rel ** 2).mean())),
            "max_abs": float(np.abs(a - b).max()), "n_diff": int((a != b).sum()), "n": int(a.size)}
