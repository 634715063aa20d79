"""BASELINE.json configs[4] as a pipeline over the C ABI: PTO-style stitch of 6 rectilinear positions x 3 exposure
brackets -> spherical panorama, the way the reference can run it:

    stage A   per position, `--synopsis hdr_merge --single 0` of its three brackets      (6 jobs)
    stage B   `--synopsis panorama` (voronoi) over the six merged rasters -> 16384 x 8192

and its partition over the GPUs of one box (reference work splitting: zimt/wielding.h:251-265
hands out lines x segments of ONE frame): contiguous ROW BANDS of the panorama, one per rank, sized by
estimated cost. Nothing is exchanged between the stages: a rank's band samples a curved region of every
position, so the rank uploads just the rectangles of the 18 bracket rasters that cover that region
(eu_source_reserve + eu_source_write_rect), merges exactly those rectangles itself - rendered straight into the
container of stage B's source (eu_render_rect_pitched, eu_source_commit) - and stitches its band. The band goes
to the host over the rank's own PCIe link, into ONE frame in shared page-locked memory. The only collectives
are barriers and the reductions of timings and checksums.

Which part of a merged raster does stage B read? The winner of _voronoi_syn (envutil_payload.cc:818-956) is
the facet with the largest z * recip_step among those the ray hits; the six positions share one step and differ
in yaw only (pitch = roll = 0), so z = cos(dlon) cos(lat) picks the position nearest in longitude: a position
is only ever EVALUATED for |dlon| <= 30 degrees (plus rounding: a margin of a few texels), i.e. for image-plane
u = tan(dlon) in +-0.577 of +-1.19 - 49 % of its columns - and, for a band of latitudes, rows
v = tan(lat) sqrt(1 + u^2). A facet further away in longitude that the ray also hits never wins. The final
panorama is bit-identical to the one made from fully merged rasters (bench.py and tests/ check that); what is
skipped is dead work, and the bench line says so (config.c5.plan).

Host logic (the plan) is plain numpy and runs without a GPU; the pipeline class needs torch + the library.
"""
import math

import numpy as np

from . import bands as eu_bands

POSITIONS, BRACKETS = 6, 3
HFOV_DEG, YAW_STEP_DEG = 100.0, 60.0
EVS = (12.0, 10.0, 14.0)  # Eev of the brackets, middle exposure first (workloads.C5_BRACKETS)
COL_MARGIN, ROW_MARGIN = 4, 3  # texels around the analytic region: bilinear window (1) + float rounding of the rays
N_STRIPS = 4  # column strips per position: narrower strips follow the curved region more closely


def sizes(scale=1):
    """(w, h) of a position's raster and (W, H) of the panorama."""
    return (6000 // scale, 4000 // scale), (16384 // scale, 8192 // scale)


def _extent(w, h):
    ex = math.tan(math.radians(HFOV_DEG) / 2.0)
    return ex, ex * h / w


def col_of(u, w, h):
    ex, _ = _extent(w, h)
    return (u / (2.0 * ex) + 0.5) * w - 0.5


def row_of(v, w, h):
    _, ey = _extent(w, h)
    return (v / (2.0 * ey) + 0.5) * h - 0.5


def needed_columns(w, h):
    """Columns [c0, c1) of a position that stage B can evaluate (|dlon| <= half the yaw step), c0 a multiple of 32."""
    u = math.tan(math.radians(YAW_STEP_DEG) / 2.0)
    c0 = int(math.floor(col_of(-u, w, h))) - COL_MARGIN
    c1 = int(math.ceil(col_of(u, w, h))) + 2 + COL_MARGIN
    c0 = max(0, (c0 // 32) * 32)
    return c0, min(w, c1)


def strip_edges(w, h, n_strips=N_STRIPS):
    c0, c1 = needed_columns(w, h)
    n = max(1, min(n_strips, (c1 - c0) // 64))
    edges = [c0 + ((c1 - c0) * k // n) // 32 * 32 for k in range(n)] + [c1]
    return sorted(set(edges))


def lat_of_row(r, H):
    """Latitude of the CENTRE of panorama row r (edge-to-edge pixels, stepper.h:324-333)."""
    return ((r + 0.5) / H - 0.5) * math.pi


def rects_for_band(row0, row1, w, h, H, n_strips=N_STRIPS):
    """Rectangles (r0, r1, c0, c1) of a position's raster that cover every texel the panorama rows [row0, row1)
    can sample from it. Empty when the band lies beyond the position's vertical field of view."""
    ex, ey = _extent(w, h)
    la, lb = lat_of_row(row0, H), lat_of_row(row1 - 1, H)
    lim = math.radians(89.99)
    ta, tb = math.tan(max(-lim, min(lim, la))), math.tan(max(-lim, min(lim, lb)))
    edges = strip_edges(w, h, n_strips)
    out = []
    for c0, c1 in zip(edges[:-1], edges[1:]):
        # image-plane u of the strip's texel centres, widened by the margin
        ua = ((c0 - COL_MARGIN + 0.5) / w - 0.5) * 2.0 * ex
        ub = ((c1 + COL_MARGIN - 0.5) / w - 0.5) * 2.0 * ex
        amin = 0.0 if ua <= 0.0 <= ub else min(abs(ua), abs(ub))
        amax = max(abs(ua), abs(ub))
        smin, smax = math.sqrt(1.0 + amin * amin), math.sqrt(1.0 + amax * amax)
        vs = [ta * smin, ta * smax, tb * smin, tb * smax]
        vmin, vmax = min(vs), max(vs)
        if vmin > ey or vmax < -ey:  # the whole strip misses the image for these latitudes
            continue
        r0 = int(math.floor(row_of(max(vmin, -ey), w, h))) - ROW_MARGIN
        r1 = int(math.floor(row_of(min(vmax, ey), w, h))) + 2 + ROW_MARGIN
        r0, r1 = max(0, r0), min(h, r1)
        if r1 > r0:
            out.append((r0, r1, c0, c1))
    return out


def full_rects(w, h):
    """The whole raster as one rectangle: stage A as the reference runs it (every texel of every position merged)."""
    return [(0, h, 0, w)]


def row_costs(w, h, W, H, n_strips=N_STRIPS, c_a=28.4, c_b=18.8, c_b0=13.2):
    """Estimated cost of every panorama row in picoseconds: stage B's pixels (c_b where a position is in sight, c_b0
    elsewhere: six rays and six mask tests are computed either way) + the rows of the six positions that this row adds
    to stage A (c_a per merged texel). The constants are measured on a B200 (tools/calibrate_c5_cost.py,
    profiles/r02f_c5_cost_calibration.json)."""
    ex, ey = _extent(w, h)
    edges = strip_edges(w, h, n_strips)
    lat = np.array([lat_of_row(r, H) for r in range(H)])
    step = math.pi / H
    cost = np.empty(H)
    top = math.atan(ey)  # highest latitude any position shows (at its centre column)
    for i, la in enumerate(lat):
        if abs(la) > top + step:
            cost[i] = W * c_b0
            continue
        a_rows = 0.0
        for c0, c1 in zip(edges[:-1], edges[1:]):
            um = ((0.5 * (c0 + c1)) / w - 0.5) * 2.0 * ex
            s = math.sqrt(1.0 + um * um)
            if abs(math.tan(la)) * s > ey * 1.02:
                continue
            # d(row)/d(lat) = h / (2 ey) * sec^2(lat) * s
            a_rows += (c1 - c0) * (h / (2.0 * ey)) * s / (math.cos(la) ** 2) * step
        cost[i] = W * c_b + POSITIONS * a_rows * c_a
    return cost


def plan_bands(world, scale=1, balance="cost"):
    """[(row0, row1)] * world: the ranks' bands of the panorama."""
    (w, h), (W, H) = sizes(scale)
    if world == 1:
        return [(0, H)]
    if balance == "equal":
        return eu_bands.bands(H, world)
    return eu_bands.weighted_bands(row_costs(w, h, W, H), world)


def stage_a_pixels(rects):
    return sum((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in rects)


# ---------------------------------------------------------------------------------------------------------
def synth_rect(p, rect, w, h):
    """The three brackets of position p inside `rect` = (r0, r1, c0, c1), exactly the texels workloads.c5_facets
    would put there (the scene is a function of the texel's direction and index)."""
    from . import synth
    r0, r1, c0, c1 = rect
    base = synth.rectilinear_facet(w, h, HFOV_DEG, YAW_STEP_DEG * p, 0.0, 0.0, rows=(r0, r1), cols=(c0, c1))
    return [np.clip(base * np.float32(2.0 ** (12.0 - ev)), 0.0, 1.0).astype(np.float32) for ev in EVS]


class Pipeline:
    """One rank's share of the C5 pipeline (rank 0 of 1 = the whole job on one GPU).

    plan: "needed" = merge only the rectangles stage B samples (default), "full" = merge every texel of every
    position as the reference's stage A does. Rasters live in page-locked host memory; step_e2e() moves them."""

    def __init__(self, engine, torch, rank=0, world=1, scale=1, plan="needed", balance="cost", contracted=None,
                 host_frame=None, synth_inputs=True, a_streams=None):
        from . import workloads
        from .job import FacetSpec
        self.eng, self.torch, self.rank, self.world, self.scale, self.plan = engine, torch, rank, world, scale, plan
        (self.w, self.h), (self.W, self.H) = sizes(scale)
        w, h = self.w, self.h
        self.bands = plan_bands(world, scale, balance)
        self.row0, self.row1 = self.bands[rank]
        self.rects = rects_for_band(self.row0, self.row1, w, h, self.H) if plan == "needed" else full_rects(w, h)
        self.stream = torch.cuda.current_stream().cuda_stream
        lib = engine.lib
        # job descriptions (geometry only): stage A per position, stage B over the merged rasters
        self.jobs_a, self.st_a = [], []
        for p in range(POSITIONS):
            fs = [FacetSpec(None, "rectilinear", HFOV_DEG, yaw=YAW_STEP_DEG * p, eev=ev, width=w, height=h, nchannels=3)
                  for ev in EVS]
            job, _ = workloads.c5_stage_a_geometry(fs, w, h)
            job.contracted = contracted
            self.jobs_a.append(job)
            self.st_a.append(job.structs(lib))
        fsb = [FacetSpec(None, "rectilinear", HFOV_DEG, yaw=YAW_STEP_DEG * p, width=w, height=h, nchannels=3)
               for p in range(POSITIONS)]
        self.job_b, self.alg_b_full = workloads.c5_stage_b_geometry(fsb, scale)
        self.job_b.contracted = contracted
        self.st_b = self.job_b.structs(lib)
        # containers: 18 bracket sources + 6 merged sources, written in place
        from . import capi
        self.src_a = []  # [p][b] -> handle
        for p in range(POSITIONS):
            t, fa, o, taps, ntaps = self.st_a[p]
            self.src_a.append([engine.reserve(fa[b], o) for b in range(BRACKETS)])
        self.src_b = [engine.reserve(self.st_b[1][p], self.st_b[2]) for p in range(POSITIONS)]
        self.hs_a = []
        for p in range(POSITIONS):
            hs = (capi.SourceH * BRACKETS)()
            for b in range(BRACKETS):
                hs[b] = self.src_a[p][b][0]
            self.hs_a.append(hs)
        self.hs_b = (capi.SourceH * POSITIONS)()
        for p in range(POSITIONS):
            self.hs_b[p] = self.src_b[p][0]
        # inputs in page-locked host memory, one block per (position, bracket, rectangle)
        self.host = {}
        self.h2d_bytes = 0
        if synth_inputs:
            for p in range(POSITIONS):
                for k, rect in enumerate(self.rects):
                    for b, img in enumerate(synth_rect(p, rect, w, h)):
                        t = torch.from_numpy(img).pin_memory()
                        self.host[(p, b, k)] = t
                        self.h2d_bytes += t.numel() * 4
        nb = self.row1 - self.row0
        self.d_band = [torch.empty((nb, self.W, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
        self.copy_stream = torch.cuda.Stream()
        # stage A on several streams (see stage_a)
        n_side = a_streams if a_streams is not None else 4
        self.side_streams = [torch.cuda.Stream() for _ in range(n_side)]
        self.band_done = [None, None]  # events: the D2H of band buffer k has finished
        self.host_frame = host_frame   # H x W x 3 float32 tensor in page-locked (shared) host memory, or None
        self.d2h_bytes = nb * self.W * 3 * 4
        self.step_no = 0

    # ---- the stages ----------------------------------------------------------------------------------
    def upload(self):
        """H2D of this rank's part of the 18 bracket rasters + their braces."""
        eng = self.eng
        for p in range(POSITIONS):
            for b in range(BRACKETS):
                hnd = self.src_a[p][b][0]
                for k, (r0, r1, c0, c1) in enumerate(self.rects):
                    t = self.host[(p, b, k)]
                    eng.write_rect(hnd, t.data_ptr(), (c1 - c0) * 3, r0, r1, c0, c1, self.stream)
                eng.commit(hnd, self.st_a[p][1][b], self.st_a[p][2], self.stream, timed=False)

    def _fork(self):
        main = self.torch.cuda.current_stream()
        fork = self.torch.cuda.Event()
        fork.record(main)
        for s in self.side_streams:
            s.wait_event(fork)
        return main

    def _join(self, main):
        for s in self.side_streams:
            join = self.torch.cuda.Event()
            join.record(s)
            main.wait_event(join)

    def stage_a(self):
        """The merges of the rectangles: independent launches (disjoint parts of six containers). Back to back on one
        stream every launch pays the drain of the one before it - 8 us each, 0.19 ms per step whatever the band
        (tools/probe_c5_rank.py, profiles/r02l_rank_probe.txt: a rank of 8 spends 0.54 ms here on one stream, 0.34 ms
        on two or more) - so the launches are dealt to a few streams (fork and join on the pipeline's stream)."""
        eng = self.eng
        side = self.side_streams
        main = self._fork() if side else None
        k = 0
        for p in range(POSITIONS):
            hnd, core, pitch = self.src_b[p]
            for (r0, r1, c0, c1) in self.rects:
                st = side[k % len(side)].cuda_stream if side else self.stream
                k += 1
                eng.render_rect_pitched(self.jobs_a[p], self.hs_a[p], self.st_a[p], r0, r1, c0, c1, core + r0 * pitch * 4,
                                        pitch, st, texel_floats=eng.texel_floats[hnd.value])
        if side:
            self._join(main)

    def stage_b_staging(self):
        """Brace of the six merged rasters: a few tiny launches each, one position per stream."""
        side = self.side_streams
        main = self._fork() if side else None
        for p in range(POSITIONS):
            st = side[p % len(side)].cuda_stream if side else self.stream
            self.eng.commit(self.src_b[p][0], self.st_b[1][p], self.st_b[2], st, timed=False)
        if side:
            self._join(main)

    def stage_b(self, k=0):
        self.eng.render_rows(self.job_b, self.hs_b, self.st_b, self.row0, self.row1, self.d_band[k].data_ptr(), self.stream,
                             timed=False)

    def step_device(self, k=0):
        """Stage A + brace + stage B on the resident brackets (the device-timed unit)."""
        self.stage_a()
        self.stage_b_staging()
        self.stage_b(k)

    def step_e2e(self):
        """Host rasters in, host band out: H2D + stage A + stage B + D2H into the shared frame. The download of step
        n runs on its own stream and overlaps the upload of step n+1 (two band buffers)."""
        torch = self.torch
        k = self.step_no & 1
        self.step_no += 1
        cur = torch.cuda.current_stream()
        if self.band_done[k] is not None:
            cur.wait_event(self.band_done[k])
        self.upload()
        self.step_device(k)
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ev)
            if self.host_frame is not None:
                self.host_frame[self.row0:self.row1].copy_(self.d_band[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.band_done[k] = done

    def finish(self):
        self.torch.cuda.current_stream().synchronize()
        self.copy_stream.synchronize()

    def stage_a_alg_bytes(self):
        """Algorithmic bytes of this rank's stage A: every merged texel is stored once and reads one texel of each
        bracket (the target has the geometry of the middle bracket: sampling lands on texel centres)."""
        return stage_a_pixels(self.rects) * POSITIONS * 12 * (1 + BRACKETS)

    def close(self):
        self.finish()
        for p in range(POSITIONS):
            self.eng.release([h for h, _, _ in self.src_a[p]])
        self.eng.release([h for h, _, _ in self.src_b])
        self.src_a, self.src_b, self.host = [], [], {}
