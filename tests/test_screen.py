"""Tethered output (to_screen_t, reference envutil_payload.cc:298-413,524-531): uint32 sRGBA frames.

The golden frames (tests/golden/screen.json) were written by the UNMODIFIED reference running tethered
(tools/make_golden_screen.py). CPU tests pin the oracle's restatement and the product's table to them; the GPU tests
compare eu_render_screen / eu_to_screen_device through the C ABI."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import harness
import jobs
from envutil_b200 import capi

MANIFEST = json.load(open(os.path.join(harness.GOLDEN, "screen.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u4").tobytes()).hexdigest()


def oracle_screen(px):
    lib = harness.oracle()
    lib.orc_to_screen.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    px = np.ascontiguousarray(px, dtype=np.float32)
    out = np.empty(px.shape[:-1], dtype=np.uint32)
    lib.orc_to_screen(px.ctypes.data, px.shape[-1], out.size, out.ctypes.data)
    return out


def oracle_frame(job):
    st = job.structs()
    st[0].gain = 0.0  # work() does not un-brighten 'single' jobs when it runs tethered (envutil_payload.cc:491)
    lib = harness.oracle()
    t, fa, o, taps, ntaps = st
    hs = harness.oracle_sources(job, st)
    oh, ow = t.out_shape()
    out = np.empty((oh, ow, t.nchannels), dtype=np.float32)
    rc = lib.orc_render(C.byref(t), C.byref(o), len(job.facets), fa, hs, taps, ntaps, 0, oh, out.ctypes.data, None, 0)
    assert rc == 0
    for h in hs:
        lib.orc_source_free(h)
    return oracle_screen(out)


def test_manifest_covers_the_screen_jobs():
    assert sorted(MANIFEST) == sorted(jobs.SCREEN_JOBS)


@pytest.mark.parametrize("name", jobs.SCREEN_JOBS)
def test_oracle_screen_equals_golden(name):
    got = oracle_frame(jobs.JOBS[name])
    assert list(got.shape) == MANIFEST[name]["shape"]
    assert _sha(got) == MANIFEST[name]["sha256"]


@pytest.mark.parametrize("name", ["ll_rect_d1", "rgba1_rect_d1"])
def test_oracle_screen_equals_golden_frame(name):
    gold = np.load(os.path.join(harness.GOLDEN, "screen_" + name + ".npz"))["out"]
    assert np.array_equal(oracle_frame(jobs.JOBS[name]), gold)


@pytest.mark.skipif(harness.ref_binary("pm") is None, reason="oracle/_ref not built (only where /root/reference exists)")
def test_oracle_screen_equals_live_reference():
    for name in ("ga_cm_sph_d3", "ll_fish_d1_tw4"):
        assert np.array_equal(oracle_frame(jobs.JOBS[name]), harness.reference_screen(jobs.JOBS[name], "pm")), name


def test_product_table_equals_oracle_table_and_known_answers():
    lib = capi.load()
    mine, theirs = (C.c_float * 257)(), (C.c_float * 257)()
    lib.eu_screen_lut(mine)
    orc = harness.oracle()
    orc.orc_screen_lut.argtypes = [C.POINTER(C.c_float)]
    orc.orc_screen_lut(theirs)
    a, b = np.frombuffer(mine, dtype=np.float32), np.frombuffer(theirs, dtype=np.float32)
    assert a.tobytes() == b.tobytes()
    assert a[0] == 0.0 and a[255] == 255.0 and np.all(np.diff(a[:256]) > 0)
    # sRGB of linear 0.5 is 0.7354 (IEC 61966-2-1): knot 128 = 0.50196 -> 188.0 within the table's resolution
    assert abs(a[128] - 255.0 * (1.055 * (128 / 255.0) ** (1 / 2.4) - 0.055)) < 1e-3
    # the packing: mid grey, opaque; values beyond [0, 1] clamp
    px = np.array([[0.5, 0.5, 0.5], [-1.0, 2.0, 1.0], [0.0, 0.0031308, 1e-9]], dtype=np.float32)
    out = oracle_screen(px)
    assert out[0] == 0xFF000000 | (187 << 16) | (187 << 8) | 187
    assert out[1] == 0xFF000000 | (255 << 16) | (255 << 8) | 0
    assert out[2] == 0xFF000000 | (0 << 16) | (10 << 8) | 0


# ---- GPU ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", jobs.SCREEN_JOBS)
def test_gpu_screen_equals_golden(engine, name):
    got = engine.render_screen(jobs.JOBS[name])
    assert list(got.shape) == MANIFEST[name]["shape"]
    assert _sha(got) == MANIFEST[name]["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("nch", [1, 2, 3, 4])
def test_gpu_to_screen_device_equals_oracle(engine, nch):
    import torch
    rng = np.random.default_rng(40 + nch)
    px = rng.uniform(-0.2, 1.2, size=(257, 193, nch)).astype(np.float32)
    px[0, :8] = np.array([0.0, 1.0, 0.0031308, 0.5, 1.0 / 255.0, 254.5 / 255.0, np.nextafter(np.float32(1), np.float32(0)), 1e-30],
                         dtype=np.float32)[:, None]
    d = torch.from_numpy(px).cuda()
    out = torch.empty(px.shape[:2], dtype=torch.int32, device="cuda")
    capi.check(engine.lib.eu_to_screen_device(C.c_void_p(d.data_ptr()), nch, px.shape[0] * px.shape[1],
                                              C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               engine.lib)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, oracle_screen(px))
