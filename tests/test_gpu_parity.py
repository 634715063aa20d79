"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the
same inputs. Bar (BASELINE.json north_star): pixel values within 1e-5 relative, face / facet
indices bit-exact. Because the kernels use the same binary32 elementary functions and the
same operation order as the oracle (include/eu_math.h, -fmad=false), the tests demand more:
every float of every small job is BIT-IDENTICAL (0 differing values)."""
import copy
import os
import ctypes as C

import numpy as np
import pytest

import harness
import jobs
from envutil_b200 import capi

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: relative, on linear float RGB, denominator floored at 1e-3


@pytest.mark.parametrize("name", sorted(jobs.JOBS))
def test_small_job_bit_exact(engine, name):
    job = jobs.JOBS[name]
    out = engine.render(job)
    ref, ridx = harness.oracle_render(job, want_index=True)
    c = harness.compare(out, ref)
    assert c["max_rel"] <= TOL, c
    assert c["n_diff"] == 0, c
    if job.twine == 0:
        idx = engine.index_plane(job)
        assert np.array_equal(idx, ridx)


@pytest.mark.parametrize("name", ["ll_rect_d3_rot", "ll_rect_d5", "ll_rect_d2", "llpart_rect_d3", "ll360x90_rect_d3",
                                  "cyl360_src_sph_d2", "cm_sph_d3", "cm100_sph_d2", "cm_sph_d5", "ba6_sph_d1",
                                  "cm_sph_d1_support4_tile16", "rect_src_rect_d3", "ll_rect_d7"])
def test_staged_container_bit_exact(engine, name):
    """Staging: brace, b-spline prefilter (all boundary conditions), spherical over-the-pole
    prefilter, cubemap IR with support fill - the coefficient container equals the oracle's."""
    job = jobs.JOBS[name]
    st = job.structs()
    hs = engine.stage(job, st)
    ohs = harness.oracle_sources(job, st)
    try:
        for h, oh in zip(hs, ohs):
            got, shp = engine.container(h)
            p, oshp = harness.oracle_container(oh)
            assert shp == oshp
            want = np.ctypeslib.as_array(p, shape=(got.size,))
            assert np.array_equal(got, want), (name, int((got != want).sum()), got.size)
    finally:
        engine.release(hs)
        for oh in ohs:
            harness.oracle().orc_source_free(oh)


@pytest.mark.parametrize("name", ["ll_rect_d1_rot", "cm_sph_d3", "voronoi4_sph_d1", "ll_fish_d1_tw4"])
@pytest.mark.parametrize("padded", [True, False])
def test_texel_layouts_are_bit_exact(engine, name, padded):
    """The 16-byte texel layout (one 128-bit load per tap; the library's default for bilinear RGB jobs) and the
    12-byte one (its default for higher degrees) change addresses, not values: both forced, for every listed job."""
    job = copy.copy(jobs.JOBS[name])
    job.padded = padded
    assert harness.compare(engine.render(job), harness.oracle_render(job))["n_diff"] == 0


# job -> the compiled-in job shape (plan.h: eu_render_specs) whose kernel must have rendered it
SHAPED = {"cm_sph_d1": 1, "cm_sph_d3": 1, "cm_sph_d3_rot": 1, "cm_sph_d1_support4_tile16": 1, "ba6_sph_d1": 2,
          "ll_rect_d1": 3, "ll_rect_d1_rot": 3, "ll_rect_d3_rot": 3, "ll_rect_d1_oddangles": 3, "ll_ba6_d1": 4,
          "ll_fish_d1_tw4": 5, "hdr3_rect_d1": 6, "voronoi4_sph_d1": 7, "win_voronoi_sph_d1": 7,
          # same projections, but outside what the shape kernels were compiled for -> general kernels
          "ba6_sph_d3_rot": 0, "cm100_sph_d2": 0, "ll_fish_d3_rot": 0, "lens3_voronoi_sph_d1": 0, "llpart_rect_d3": 0,
          "voronoi4_sph_d3_rot": 0, "grey_cm_sph_d1_tw2": 0, "rgba_cm_sph_d1": 0, "tr1_sph_d1": 0}


@pytest.mark.parametrize("name", sorted(SHAPED))
def test_shape_kernels_and_general_kernels_agree(engine, name):
    """Jobs of the commonest shapes run kernels with the target projection / source kind / gates
    compiled in (render_spec.cu); no_spec forces the general kernels. Both equal the oracle bit for
    bit, and the library reports which one ran."""
    ref = harness.oracle_render(jobs.JOBS[name])
    for no_spec in (False, True):
        job = copy.copy(jobs.JOBS[name])
        job.no_spec = no_spec
        out = engine.render(job)
        assert engine.last_timing.shape == (0 if no_spec else SHAPED[name]), (name, no_spec, engine.last_timing.shape)
        assert harness.compare(out, ref)["n_diff"] == 0, (name, no_spec)


@pytest.mark.parametrize("name", ["ll_rect_d1_rot", "ll_rect_d3_rot", "cm_sph_d3", "ba6_sph_d1", "ll_fish_d1_tw4",
                                  "ll_cube_d1", "cm_rect_d1_tw3_rot", "lens1_rect_d3_tw2", "grey_ll_rect_d3"])
@pytest.mark.parametrize("padded", [False, True])
def test_direct_and_staged_kernels_agree(engine, name, padded):
    """Single-facet jobs run the footprint-staged kernel (cp.async.bulk into shared memory) by
    default; no_tiles forces the direct-gather kernel. Both must equal the oracle bit for bit."""
    ref = harness.oracle_render(jobs.JOBS[name])
    for no_tiles in (False, True):
        job = copy.copy(jobs.JOBS[name])
        job.padded, job.no_tiles = padded, no_tiles
        out = engine.render(job)
        c = harness.compare(out, ref)
        assert c["n_diff"] == 0, (no_tiles, c)


def test_twine_single_centre_tap_equals_plain(engine):
    """twining with one tap (0,0,1) == no twining (normalised rays differ from
    unnormalised ones only by scale, which lat/lon lookup ignores up to rounding)."""
    job = copy.deepcopy(jobs.JOBS["ll_rect_d1_rot"])
    st = list(job.structs())
    plain = engine.render(job, structs=tuple(st))
    taps = (capi.Tap * 1024)()
    taps[0].x, taps[0].y, taps[0].w = 0.0, 0.0, 1.0
    st[3], st[4] = taps, 1
    tw = engine.render(job, structs=tuple(st))
    c = harness.compare(tw, plain)
    assert c["max_rel"] < 1e-4, c


def test_row_bands_tile_the_full_render(engine):
    """eu_render_rows (the multi-GPU work unit): bands rendered separately equal the full frame."""
    torch = pytest.importorskip("torch")
    job = jobs.JOBS["voronoi4_sph_d3_rot"]
    st = job.structs()
    t = st[0]
    hs = engine.stage(job, st)
    try:
        full = engine.render(job, sources=hs, structs=st)
        buf = torch.empty((t.height, t.width, t.nchannels), dtype=torch.float32, device="cuda:0")
        edges = [0, 7, 40, 41, t.height]
        stream = torch.cuda.current_stream().cuda_stream
        for r0, r1 in zip(edges[:-1], edges[1:]):
            engine.render_rows(job, hs, st, r0, r1, buf[r0].data_ptr(), stream)
        torch.cuda.synchronize()
        assert np.array_equal(buf.cpu().numpy(), full)
    finally:
        engine.release(hs)


def test_asset_cache_cycles(engine):
    """Two-generation cache keyed by asset_key (environment.h:84-227): an asset survives the
    cycle it was used in, and is dropped after a cycle in which it was not used."""
    lib = engine.lib
    job = jobs.JOBS["ll_rect_d1"]
    st = job.structs()
    hs = engine.stage(job, st, keys=["test-asset-A"])
    assert lib.eu_source_find(b"test-asset-A") == hs[0]
    assert lib.eu_cycle() == 0
    assert lib.eu_source_find(b"test-asset-A") == hs[0]  # touched in the new cycle
    assert lib.eu_cycle() == 0
    assert lib.eu_cycle() == 0  # not used during the cycle that just ended -> dropped
    assert lib.eu_source_find(b"test-asset-A") is None
    assert lib.eu_source_release(hs[0]) != 0  # the handle is dead


def test_errors_are_reported_not_fatal(engine):
    lib = engine.lib
    job = copy.deepcopy(jobs.JOBS["ll_rect_d1"])
    st = job.structs()
    hs = engine.stage(job, st)
    try:
        t, fa, o, taps, ntaps = st
        out = np.empty((t.height, t.width, 3), dtype=np.float32)
        o2 = capi.Opts.from_buffer_copy(o)
        o2.spline_degree = 3  # staged for degree 1
        rc = lib.eu_render(C.byref(t), C.byref(o2), 1, fa, hs, taps, 0, out.ctypes.data, None)
        assert rc == -1 and b"degree" in lib.eu_last_error()
        bogus = (capi.SourceH * 1)(C.c_void_p(0x1234))
        assert lib.eu_render(C.byref(t), C.byref(o), 1, fa, bogus, taps, 0, out.ctypes.data, None) == -1
        # nonsense is refused, never approximated: a window larger than the image
        crop = capi.Facet.from_buffer_copy(fa[0])
        crop.window_width = crop.width * 2
        h = capi.SourceH()
        img = np.ascontiguousarray(job.facets[0].image)
        assert lib.eu_source_upload(None, C.byref(crop), C.byref(o), img.ctypes.data, C.byref(h), None) == -1
        assert b"does not lie inside" in lib.eu_last_error()
    finally:
        engine.release(hs)


def test_async_jobs_do_not_share_plan_buffers(engine):
    """eu_render_rows is asynchronous on the caller's stream. Different multi-facet jobs enqueued
    back to back must not see each other's facet array / tap list (they live in one device buffer
    each and are rewritten in stream order) - found with the multi-GPU C5 pipeline."""
    torch = pytest.importorskip("torch")
    names = ["voronoi4_sph_d1", "hdr3_rect_d1", "voronoi3_rect_d1_tw2", "hdr3_sph_d3_tw2", "voronoi4_sph_d3_rot"]
    stream = torch.cuda.current_stream().cuda_stream
    work = []
    for n in names:
        job = jobs.JOBS[n]
        st = job.structs()
        hs = engine.stage(job, st)
        t = st[0]
        buf = torch.empty((t.height, t.width, t.nchannels), dtype=torch.float32, device="cuda:0")
        work.append((n, job, st, hs, buf))
    try:
        for _ in range(3):
            for n, job, st, hs, buf in work:
                engine.render_rows(job, hs, st, 0, st[0].height, buf.data_ptr(), stream, timed=False)
        torch.cuda.synchronize()
        for n, job, st, hs, buf in work:
            assert np.array_equal(buf.cpu().numpy(), harness.oracle_render(job)), n
    finally:
        for n, job, st, hs, buf in work:
            engine.release(hs)


def test_pipelined_jobs_match_blocking_calls(engine):
    """eu_source_upload_async / eu_render_async / eu_job_wait: several jobs in flight on the
    upload, render and download streams give the frames the blocking calls give."""
    torch = pytest.importorskip("torch")
    names = ["cm_sph_d3", "ll_rect_d3_rot", "ll_fish_d1_tw4", "ba6_sph_d1", "cm_sph_d3_rot", "ll_cube_d1"]
    tickets, outs = [], []
    for rep in range(2):
        for n in names:
            job = jobs.JOBS[n]
            st = job.structs()
            src = torch.from_numpy(np.ascontiguousarray(job.facets[0].image)).pin_memory()
            out = torch.empty((st[0].height, st[0].width, st[0].nchannels), dtype=torch.float32).pin_memory()
            tickets.append((engine.submit(job, st, [src.data_ptr()], out.data_ptr()), src))
            outs.append((n, out))
            if len(tickets) >= 3:
                tk, _ = tickets.pop(0)
                tm = engine.finish(tk)
                assert tm.render_ms > 0
    while tickets:
        engine.finish(tickets.pop(0)[0])
    for n, out in outs:
        assert np.array_equal(out.numpy(), harness.oracle_render(jobs.JOBS[n])), n
    # a fifth pending job is refused
    job = jobs.JOBS["ll_rect_d1"]
    st = job.structs()
    src = torch.from_numpy(np.ascontiguousarray(job.facets[0].image)).pin_memory()
    bufs = [torch.empty((st[0].height, st[0].width, 3), dtype=torch.float32).pin_memory() for _ in range(5)]
    tks = [engine.submit(job, st, [src.data_ptr()], bufs[i].data_ptr()) for i in range(4)]
    with pytest.raises(RuntimeError, match="in flight"):
        engine.submit(job, st, [src.data_ptr()], bufs[4].data_ptr())
    for tk in tks:
        engine.finish(tk)
    engine.lib.eu_cycle()


@pytest.mark.parametrize("degree", [1, 3])
def test_stage_one_renders_into_stage_two_source(engine, degree):
    """Two-stage jobs (BASELINE configs[4]): the hdr_merge of a position's brackets is rendered straight
    into the container of the source the panorama stage reads (eu_source_reserve +
    eu_render_rows_pitched + eu_source_commit). The container equals the one obtained by rendering
    into a dense raster and uploading that, and so does a panorama rendered from it."""
    import ctypes as C
    import torch
    from envutil_b200.job import FacetSpec, Job
    ja = copy.copy(jobs.JOBS["single0_hdr3_d1"])          # stage one: --synopsis hdr_merge --single 0
    sta = ja.structs(engine.lib)
    ta = sta[0]
    hsa = engine.stage(ja, sta)
    dense = torch.empty((ta.height, ta.width, ta.nchannels), dtype=torch.float32, device="cuda")
    engine.render_rows(ja, hsa, sta, 0, ta.height, dense.data_ptr(), 0, timed=True)
    f0 = ja.facets[0]
    jb = Job([FacetSpec(None, f0.projection, f0.hfov, yaw=f0.yaw, pitch=f0.pitch, roll=f0.roll, width=ta.width,
                        height=ta.height, nchannels=ta.nchannels)], "spherical", 360.0, 256, 128, degree=degree)
    stb = jb.structs(engine.lib)
    fb, ob = stb[1], stb[2]
    via_upload = engine.stage_device(jb, [dense.data_ptr()], stb)
    h, core, pitch = engine.reserve(fb[0], ob)
    try:
        # two bands, to exercise the row offset inside the container
        mid = ta.height // 3
        tex = engine.texel_floats[h.value]  # 4 for the bilinear panorama (16-byte RGB texels), 3 for the cubic one
        assert tex == (4 if degree == 1 else 3)
        engine.render_rows_pitched(ja, hsa, sta, 0, mid, core, pitch, 0, timed=True, texel_floats=tex)
        engine.render_rows_pitched(ja, hsa, sta, mid, ta.height, core + mid * pitch * 4, pitch, 0, timed=True, texel_floats=tex)
        engine.commit(h, fb[0], ob)
        got, shp = engine.container(h)
        want, wshp = engine.container(via_upload[0])
        assert shp == wshp
        assert np.array_equal(got, want)  # the download strips the row padding
        one = (type(via_upload))(h)
        out_a = engine.render(jb, sources=one, structs=stb)
        out_b = engine.render(jb, sources=via_upload, structs=stb)
        assert np.array_equal(out_a, out_b)
    finally:
        engine.release([h])
        engine.release(via_upload)
        engine.release(hsa)


@pytest.mark.parametrize("degree,source", [(1, "host"), (1, "device"), (3, "host"), (3, "device")])
def test_write_rect_fills_a_reserved_source(engine, degree, source):
    """eu_source_write_rect: a reserved source filled piecewise - pitched and contiguous host rectangles, device
    rectangles, more rectangles than the library's scratch ring has slots, growing sizes - equals the source staged
    from the whole raster (degree 1: 16-byte texels, the rectangles pass through the scratch ring and the widening
    kernel; degree 3: 12-byte texels, pitched copies straight into the container)."""
    import torch
    from envutil_b200.job import FacetSpec, Job
    rng = np.random.default_rng(5 + degree)
    w, h = 203, 97
    img = rng.uniform(0.0, 1.0, size=(h, w, 3)).astype(np.float32)
    job = Job([FacetSpec(img, "rectilinear", 70.0, yaw=10.0)], "spherical", 360.0, 128, 64, degree=degree)
    st = job.structs(engine.lib)
    fa, o = st[1], st[2]
    whole = engine.stage(job, st)
    hnd, core, pitch = engine.reserve(fa[0], o)
    stream = torch.cuda.current_stream().cuda_stream
    keep = []
    try:
        assert engine.texel_floats[hnd.value] == (4 if degree == 1 else 3)
        # a ragged tiling: column strips of different widths, each cut into bands of growing height
        cols = [0, 1, 40, 41, 130, w]
        for c0, c1 in zip(cols[:-1], cols[1:]):
            r0, step = 0, 3
            while r0 < h:
                r1 = min(h, r0 + step)
                step += 7
                if (r0 // 3 + c0) % 2:  # pitched: a view into the whole raster
                    sub = torch.from_numpy(img).pin_memory()
                    if source == "device":
                        sub = sub.cuda()
                    ptr = sub.data_ptr() + (r0 * w + c0) * 3 * 4
                    keep.append(sub)
                    engine.write_rect(hnd, ptr, w * 3, r0, r1, c0, c1, stream)
                else:  # contiguous: its own buffer
                    sub = torch.from_numpy(np.ascontiguousarray(img[r0:r1, c0:c1])).pin_memory()
                    if source == "device":
                        sub = sub.cuda()
                    keep.append(sub)
                    engine.write_rect(hnd, sub.data_ptr(), (c1 - c0) * 3, r0, r1, c0, c1, stream)
                r0 = r1
        engine.commit(hnd, fa[0], o, stream)
        got, shp = engine.container(hnd)
        want, wshp = engine.container(whole[0])
        assert shp == wshp
        assert np.array_equal(got, want)
    finally:
        engine.release([hnd])
        engine.release(whole)


def test_edge_jobs():
    """Degenerate sizes (one-pixel and ragged targets, sources smaller than a spline window, single rows and
    columns, 4-px cube faces): the oracle equals the reference on them (golden); the kernels equal the oracle,
    or the library refuses the job. All of them run in ONE separate process with a two-minute limit, so
    that a device fault on such a size cannot reach the other tests."""
    import subprocess
    import sys
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import harness, jobs\n"
            "from envutil_b200.engine import Engine\n"
            "eng = Engine(0)\n"
            "bad = 0\n"
            "todo = dict(jobs.EDGE_JOBS, **jobs.ODD_CUBE_JOBS)\n"
            "for name in sorted(todo):\n"
            "    job = todo[name]\n"
            "    try:\n"
            "        out = eng.render(job)\n"
            "    except RuntimeError as e:\n"
            "        print(name, 'refused:', str(e)[:100], flush=True)\n"
            "        bad += 'status -2' not in str(e)\n"
            "        refused = locals().get('refused', 0) + 1\n"
            "        continue\n"
            "    c = harness.compare(out, harness.oracle_render(job))\n"
            "    print(name, c, flush=True)\n"
            "    bad += c['n_diff'] != 0\n"
            "print('refused', locals().get('refused', 0))\n"
            "sys.exit(3 if bad or locals().get('refused', 0) > 1 else 0)\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1500:]


def _rects(h, w):
    """A partition of an h x w raster into rectangles whose column edges are multiples of 32 and whose row edges
    are not multiples of the 8-row tile."""
    cols = sorted({0, w} | {c for c in (32, 96, 32 * ((w // 2) // 32 + 1)) if 0 < c < w})
    rows = sorted({0, h} | {r for r in (5, 13, h // 2 + 3) if 0 < r < h})
    return [(r0, r1, c0, c1) for r0, r1 in zip(rows[:-1], rows[1:]) for c0, c1 in zip(cols[:-1], cols[1:])]


@pytest.mark.parametrize("name", ["cm_sph_d3", "ll_rect_d1", "ll_fish_d1_tw4", "voronoi4_sph_d3_rot", "hdr3_rect_d1",
                                  "rgba4_voronoi_sph_d3", "tr1_sph_d1"])
def test_rectangles_assemble_the_frame(engine, name):
    """eu_render_rect_pitched (rows x columns, the unit of pipelines that need only part of an intermediate
    raster): a frame rendered rectangle by rectangle equals the oracle's frame bit for bit - footprint-staged
    cubic kernel, direct kernels, twining, the synopses (voronoi_plus votes per 16-lane vector), the generic stepper."""
    import torch
    job = jobs.JOBS[name]
    st = job.structs()
    t = st[0]
    h, w = t.out_shape()
    hs = engine.stage(job, st)
    try:
        buf = torch.full((h, w, t.nchannels), -7.0, dtype=torch.float32, device="cuda:0")
        stream = torch.cuda.current_stream().cuda_stream
        for r0, r1, c0, c1 in _rects(h, w):
            engine.render_rect_pitched(job, hs, st, r0, r1, c0, c1, buf[r0].data_ptr(), w * t.nchannels, stream)
        torch.cuda.synchronize()
        got = buf.cpu().numpy()
    finally:
        engine.release(hs)
    assert harness.compare(got, harness.oracle_render(job))["n_diff"] == 0


def test_rectangle_arguments_are_checked(engine):
    job = jobs.JOBS["ll_rect_d1"]
    st = job.structs()
    hs = engine.stage(job, st)
    try:
        import torch
        t = st[0]
        buf = torch.zeros((t.height, t.width, t.nchannels), dtype=torch.float32, device="cuda:0")
        for bad in ((0, 4, 16, 64), (0, 4, 0, t.width + 1), (0, 4, 64, 64), (4, 4, 0, 32)):
            with pytest.raises(RuntimeError):
                engine.render_rect_pitched(job, hs, st, bad[0], bad[1], bad[2], bad[3], buf.data_ptr(),
                                           t.width * t.nchannels, 0)
    finally:
        engine.release(hs)
