"""Thin Python driver over the C ABI (include/envutil_b200.h) for the tests, the bench and the
smoke test: stage the facets of a Job, render, read index planes. All compute happens in
libenvutil_b200.so on the GPU; there is no CPU path here and nothing imports the oracle.
"""
import ctypes as C

import numpy as np

from . import capi


class Engine:
    def __init__(self, device=0):
        self.lib = capi.load()
        capi.check(self.lib.eu_init(device), self.lib)
        self.last_timing = capi.Timing()
        self.last_stage_timing = []
        self.launches = 0  # kernels launched through this engine (claimed by bench.py)

    def close(self):
        self.lib.eu_shutdown()

    # ---- staging --------------------------------------------------------------------------
    def stage(self, job, structs=None, keys=None, padded=False):
        """Upload + brace + prefilter every facet of the job. Returns the handle array."""
        t, fa, o, taps, ntaps = structs or job.structs(self.lib)
        if padded:
            o.reserved[0] = 1
        hs = (capi.SourceH * len(job.facets))()
        self.last_stage_timing = []
        for i, f in enumerate(job.facets):
            img = np.ascontiguousarray(f.image, dtype=np.float32)
            tm = capi.Timing()
            key = keys[i].encode() if keys else None
            h = capi.SourceH()
            if f.has_alpha_spec():  # PTO exclude masks / lens crop
                a = f.alpha_spec()
                capi.check(self.lib.eu_source_upload_alpha(key, C.byref(fa[i]), C.byref(o), img.ctypes.data,
                                                           C.byref(a), C.byref(h), C.byref(tm)), self.lib)
            else:
                capi.check(self.lib.eu_source_upload(key, C.byref(fa[i]), C.byref(o), img.ctypes.data, C.byref(h),
                                                     C.byref(tm)), self.lib)
            hs[i] = h
            self.last_stage_timing.append(tm)
            self.launches += tm.launches
        return hs

    def stage_device(self, job, dev_ptrs, structs=None, stream=0, padded=False):
        """Same, rasters already in device memory (dev_ptrs: one device address per facet)."""
        t, fa, o, taps, ntaps = structs or job.structs(self.lib)
        if padded:
            o.reserved[0] = 1
        hs = (capi.SourceH * len(job.facets))()
        self.last_stage_timing = []
        for i in range(len(job.facets)):
            tm = capi.Timing()
            h = capi.SourceH()
            capi.check(self.lib.eu_source_upload_device(None, C.byref(fa[i]), C.byref(o), C.c_void_p(dev_ptrs[i]),
                                                        C.c_void_p(stream), C.byref(h), C.byref(tm)), self.lib)
            hs[i] = h
            self.last_stage_timing.append(tm)
            self.launches += tm.launches
        return hs

    # ---- sources produced on the device, in place (eu_source_reserve / eu_source_commit) -----
    def reserve(self, facet_struct, opts):
        """Container of a single-image source whose raster a render will write; returns
        (handle, device address of core texel (0,0), row pitch in floats). The texel stride in floats (the channel
        count, or 4 for RGB in the 16-byte layout) is remembered in self.texel_floats[handle]."""
        h, core, pitch, tex = capi.SourceH(), C.c_void_p(), C.c_int(), C.c_int()
        capi.check(self.lib.eu_source_reserve(None, C.byref(facet_struct), C.byref(opts), C.byref(h), C.byref(core),
                                              C.byref(pitch), C.byref(tex)), self.lib)
        if not hasattr(self, "texel_floats"):
            self.texel_floats = {}
        self.texel_floats[h.value] = tex.value
        return h, core.value, pitch.value

    def render_rows_pitched(self, job, sources, structs, row0, row1, d_out, pitch_floats, stream=0, timed=True,
                            texel_floats=0):
        t, fa, o, taps, ntaps = structs
        tm = capi.Timing()
        capi.check(self.lib.eu_render_rect_pitched(C.byref(t), C.byref(o), len(job.facets), fa, sources, taps, ntaps,
                                                   row0, row1, 0, t.out_shape()[1], C.c_void_p(d_out), pitch_floats,
                                                   texel_floats, C.c_void_p(stream), C.byref(tm) if timed else None),
                   self.lib)
        self.launches += tm.launches if timed else 1
        if timed:
            self.last_timing = tm
        return tm

    def render_rect_pitched(self, job, sources, structs, row0, row1, col0, col1, d_out, pitch_floats, stream=0,
                            texel_floats=0):
        """Columns [col0, col1) of rows [row0, row1) (col0 a multiple of 32); d_out = address of row0, column 0;
        texel_floats: floats per output pixel (0 = the channel count; 4 = RGB into a 16-byte-texel container)."""
        t, fa, o, taps, ntaps = structs
        capi.check(self.lib.eu_render_rect_pitched(C.byref(t), C.byref(o), len(job.facets), fa, sources, taps, ntaps,
                                                   row0, row1, col0, col1, C.c_void_p(d_out), pitch_floats, texel_floats,
                                                   C.c_void_p(stream), None), self.lib)
        self.launches += 1

    def commit(self, handle, facet_struct, opts, stream=0, timed=True):
        tm = capi.Timing()
        capi.check(self.lib.eu_source_commit(handle, C.byref(facet_struct), C.byref(opts), C.c_void_p(stream),
                                             C.byref(tm) if timed else None), self.lib)
        self.launches += tm.launches if timed else 1  # bilinear sources: the brace kernel
        return tm

    def write_rect(self, handle, pixels_ptr, src_pitch_floats, row0, row1, col0, col1, stream=0):
        """Rows [row0,row1) x columns [col0,col1) of a reserved source from host / device memory at pixels_ptr."""
        capi.check(self.lib.eu_source_write_rect(handle, C.c_void_p(pixels_ptr), src_pitch_floats, row0, row1, col0, col1,
                                                 C.c_void_p(stream)), self.lib)

    def release(self, handles):
        for h in handles:
            if h:
                capi.check(self.lib.eu_source_release(h), self.lib)

    def container(self, handle):
        shp = (C.c_int32 * 4)()
        n = self.lib.eu_source_container_floats(handle, shp)
        out = np.empty(n, dtype=np.float32)
        capi.check(self.lib.eu_source_download(handle, out.ctypes.data), self.lib)
        return out, tuple(shp)

    # ---- rendering ------------------------------------------------------------------------
    def render(self, job, sources=None, structs=None, out=None):
        """Render the job into a host array H x W x C (allocated unless `out` is given)."""
        st = structs or job.structs(self.lib)
        t, fa, o, taps, ntaps = st
        hs = sources if sources is not None else self.stage(job, st)
        if out is None:
            out = np.empty(t.out_shape() + (t.nchannels,), dtype=np.float32)
        tm = capi.Timing()
        try:
            capi.check(self.lib.eu_render(C.byref(t), C.byref(o), len(job.facets), fa, hs, taps, ntaps,
                                          out.ctypes.data if isinstance(out, np.ndarray) else out, C.byref(tm)),
                       self.lib)
        finally:
            if sources is None:
                self.release(hs)
        self.last_timing = tm
        self.launches += tm.launches
        return out

    def render_screen(self, job, sources=None, structs=None):
        """The job as a tethered frame: H x W uint32 sRGBA (eu_render_screen; to_screen_t of the reference)."""
        st = structs or job.structs(self.lib)
        t, fa, o, taps, ntaps = st
        hs = sources if sources is not None else self.stage(job, st)
        out = np.empty(t.out_shape(), dtype=np.uint32)
        tm = capi.Timing()
        try:
            capi.check(self.lib.eu_render_screen(C.byref(t), C.byref(o), len(job.facets), fa, hs, taps, ntaps,
                                                 out.ctypes.data, C.byref(tm)), self.lib)
        finally:
            if sources is None:
                self.release(hs)
        self.last_timing = tm
        self.launches += tm.launches
        return out

    def render_rows(self, job, sources, structs, row0, row1, d_out, stream=0, timed=True):
        """Rows [row0,row1) into device memory at address d_out, on CUDA stream `stream`."""
        t, fa, o, taps, ntaps = structs
        tm = capi.Timing()
        capi.check(self.lib.eu_render_rows(C.byref(t), C.byref(o), len(job.facets), fa, sources, taps, ntaps,
                                           row0, row1, C.c_void_p(d_out), C.c_void_p(stream),
                                           C.byref(tm) if timed else None), self.lib)
        if timed:
            self.last_timing = tm
            self.launches += tm.launches
        else:
            self.launches += 1
        return tm

    # ---- pipelined jobs (eu_source_upload_async / eu_render_async / eu_job_wait) ---------
    def submit(self, job, structs, pinned_pixels, pinned_out):
        """Enqueue upload + staging + render + download of a single-raster job; returns a ticket for
        finish(). pinned_pixels / pinned_out: addresses of page-locked host buffers."""
        t, fa, o, taps, ntaps = structs
        n = len(job.facets)
        hs = (capi.SourceH * n)()
        try:
            for i in range(n):
                h = capi.SourceH()
                capi.check(self.lib.eu_source_upload_async(None, C.byref(fa[i]), C.byref(o), C.c_void_p(pinned_pixels[i]),
                                                           C.byref(h)), self.lib)
                hs[i] = h
        except RuntimeError:  # the handles staged so far go back
            self.release(hs)
            raise
        jh = C.c_void_p()
        try:
            capi.check(self.lib.eu_render_async(C.byref(t), C.byref(o), n, fa, hs, taps, ntaps, C.c_void_p(pinned_out),
                                                C.byref(jh)), self.lib)
        except RuntimeError:
            self.release(hs)
            raise
        return jh, hs

    def finish(self, ticket):
        jh, hs = ticket
        tm = capi.Timing()
        capi.check(self.lib.eu_job_wait(jh, C.byref(tm)), self.lib)
        self.release(hs)
        self.launches += tm.launches
        return tm

    def tie_plane(self, job, sources, structs, ulps=8):
        """H x W uint8: 1 where the cube-face / winning-facet choice is within `ulps` of flipping
        (eu_debug_tie_plane); None for jobs without such a choice."""
        t, fa, o, taps, ntaps = structs
        single = len(job.facets) == 1 or o.solo >= 0
        if single and job.facets[max(o.solo, 0)].projection not in ("cubemap", "biatan6"):
            return None
        if not single and o.synopsis != capi.SYN_PANORAMA:
            return None
        tie = np.empty(t.out_shape(), dtype=np.uint8)
        capi.check(self.lib.eu_debug_tie_plane(C.byref(t), C.byref(o), len(job.facets), fa, sources, ulps,
                                               tie.ctypes.data), self.lib)
        self.launches += 1
        return tie

    def index_plane(self, job, sources=None, structs=None):
        st = structs or job.structs(self.lib)
        t, fa, o, taps, ntaps = st
        hs = sources if sources is not None else self.stage(job, st)
        idx = np.empty(t.out_shape(), dtype=np.int32)
        try:
            capi.check(self.lib.eu_debug_planes(C.byref(t), C.byref(o), len(job.facets), fa, hs, idx.ctypes.data),
                       self.lib)
        finally:
            if sources is None:
                self.release(hs)
        self.launches += 1
        return idx
