#!/usr/bin/env python3
"""Randomised pinning sweep: small random jobs through the UNMODIFIED reference build (oracle/_ref, pinned
math) and through the oracle, compared bit for bit. Only this container can run it (it needs oracle/_ref,
i.e. /root/reference); findings become jobs in tests/jobs.py with golden outputs.

  python tools/fuzz_oracle_vs_reference.py [--n 200] [--seed 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import harness  # noqa: E402
from envutil_b200.job import FacetSpec, Job  # noqa: E402

SCALE = int(os.environ.get("FUZZ_SCALE", "1"))  # multiplies the upper bounds of the random raster sizes

TARGETS = ["spherical", "cylindrical", "rectilinear", "stereographic", "fisheye", "cubemap", "biatan6"]


def random_facet(rng, nch):
    kind = rng.choice(["spherical", "cylindrical", "rectilinear", "stereographic", "fisheye", "cubemap", "biatan6"],
                      p=[.25, .08, .3, .07, .1, .12, .08])
    def img(w, h):
        a = rng.random((h, w, nch), dtype=np.float32)
        return a[:, :, 0] if nch == 1 and rng.random() < .5 else a
    if kind in ("cubemap", "biatan6"):
        f = int(rng.integers(3, 40 * SCALE))
        return FacetSpec(img(f, 6 * f), kind, float(rng.choice([90.0, 90.0, 100.0, 120.0])))
    w, h = int(rng.integers(2, 70 * SCALE)), int(rng.integers(2, 50 * SCALE))
    if kind == "spherical":
        full = rng.random() < .6
        if full:
            h = int(rng.integers(2, 30 * SCALE)); w = 2 * h
            return FacetSpec(img(w, h), kind, 360.0, yaw=float(rng.uniform(-180, 180)), pitch=float(rng.uniform(-90, 90)),
                             roll=float(rng.uniform(-30, 30)))
        return FacetSpec(img(w, h), kind, float(rng.uniform(40, 300)), yaw=float(rng.uniform(-180, 180)))
    if kind == "cylindrical":
        return FacetSpec(img(w, h), kind, float(rng.choice([360.0, rng.uniform(60, 300)])), yaw=float(rng.uniform(-180, 180)))
    if kind == "rectilinear":
        kw = {}
        if rng.random() < .35:
            kw.update(a=float(rng.uniform(-.03, .03)), b=float(rng.uniform(-.05, .05)), c=float(rng.uniform(-.02, .02)))
        if rng.random() < .2:
            kw.update(d=float(rng.uniform(-3, 3)), e=float(rng.uniform(-3, 3)))
        if rng.random() < .15:
            kw.update(g=float(rng.uniform(-2, 2)), t=float(rng.uniform(-2, 2)))
        if rng.random() < .2:
            kw.update(tr_x=float(rng.uniform(-.08, .08)), tr_y=float(rng.uniform(-.05, .05)), tr_z=float(rng.uniform(-.05, .1)))
            if rng.random() < .5:
                kw.update(tp_y=float(rng.uniform(-15, 15)), tp_p=float(rng.uniform(-10, 10)))
        if rng.random() < .25:
            kw.update(eev=float(rng.choice([10.0, 11.5, 12.0, 13.0, 14.0])))
        r = rng.random()
        if r < .12 and w > 8 and h > 8:      # lens crop (S clause) -> feathered alpha
            kw.update(crop=(int(rng.integers(0, w // 3)), int(rng.integers(2 * w // 3, w)),
                            int(rng.integers(0, h // 3)), int(rng.integers(2 * h // 3, h))))
        elif r < .24 and w > 8 and h > 8:    # exclude mask (k-line) -> feathered alpha
            n = int(rng.integers(3, 7))
            kw.update(masks=(tuple((float(rng.uniform(0, w)), float(rng.uniform(0, h))) for _ in range(n)),))
        elif r < .34:                        # the file is a window of a larger image (W clause)
            tw, th = w + int(rng.integers(0, 40)), h + int(rng.integers(0, 30))
            x0, y0 = int(rng.integers(0, tw - w + 1)), int(rng.integers(0, th - h + 1))
            kw.update(window=(x0, x0 + w, y0, y0 + h), total_width=tw, total_height=th)
        return FacetSpec(img(w, h), kind, float(rng.uniform(30, 130)), yaw=float(rng.uniform(-180, 180)),
                         pitch=float(rng.uniform(-60, 60)), roll=float(rng.uniform(-20, 20)), **kw)
    if kind == "stereographic":
        return FacetSpec(img(w, h), kind, float(rng.uniform(60, 250)), yaw=float(rng.uniform(-180, 180)),
                         pitch=float(rng.uniform(-40, 40)))
    return FacetSpec(img(max(w, h), max(w, h)), "fisheye", float(rng.choice([180.0, 360.0, rng.uniform(100, 300)])),
                     yaw=float(rng.uniform(-180, 180)))


def random_job(rng):
    nch = int(rng.choice([1, 2, 3, 3, 3, 4]))
    nf = int(rng.choice([1, 1, 1, 2, 3, 4]))
    facets = [random_facet(rng, nch) for _ in range(nf)]
    if nf > 1:  # a panorama of mounted images (cubemaps may take part)
        pass
    trg = str(rng.choice(TARGETS))
    translated = any(f.tr_x or f.tr_y or f.tr_z for f in facets)
    if translated and trg in ("cubemap", "biatan6"):
        trg = "spherical"
    width = int(rng.integers(1, 90 * SCALE))
    height = 0 if trg in ("cubemap", "biatan6") else int(rng.integers(1, 60 * SCALE))
    if trg not in ("cubemap", "biatan6") and rng.random() < .1:  # several 512-px segments per line
        width, height = int(rng.integers(513, 1200)), int(rng.integers(1, 5))
    hfov = {"spherical": float(rng.choice([360.0, 360.0, rng.uniform(40, 360)])), "cylindrical": float(rng.uniform(60, 360)), "rectilinear": float(rng.uniform(20, 140)),
            "stereographic": float(rng.uniform(60, 250)), "fisheye": float(rng.uniform(90, 300)),
            "cubemap": float(rng.choice([90.0, 100.0])), "biatan6": 90.0}[trg]
    kw = dict(degree=int(rng.choice([0, 1, 1, 1, 2, 3, 3, 4, 5, 7])), twine=int(rng.choice([0, 0, 0, 2, 3, -1])),
              yaw=float(rng.uniform(-180, 180)), pitch=float(rng.uniform(-80, 80)), roll=float(rng.uniform(-40, 40)))
    if rng.random() < .2:    # prefilter degree of its own (--prefilter)
        kw["prefilter"] = int(rng.choice([0, 1, 2, 3, 5]))
    if kw["twine"] > 0 and rng.random() < .5:
        kw.update(twine_width=float(rng.uniform(.5, 2.5)))
        if rng.random() < .5:
            kw.update(twine_sigma=float(rng.uniform(.3, 2.0)), twine_threshold=float(rng.choice([0.0, 0.01, 0.05])))
    if kw["twine"] < 0 and rng.random() < .5:
        kw.update(twine_density=float(rng.uniform(.5, 2.0)), twine_max=int(rng.integers(2, 9)))
    if any(f.projection in ("cubemap", "biatan6") for f in facets) and rng.random() < .4:
        kw.update(support_min=int(rng.choice([0, 2, 4, 8, 12])), tile_size=int(rng.choice([1, 2, 8, 16, 64])))
    if nf > 1:
        r = rng.random()
        if r < .25 and nch in (1, 3):
            kw["synopsis"] = "hdr_merge"
        elif r < .4:
            kw["solo"] = int(rng.integers(0, nf))
        elif r < .55 and not any(f.projection in ("cubemap", "biatan6") for f in facets):
            kw["single"] = int(rng.integers(0, nf))
    if trg == "spherical" and rng.random() < .5:
        height = 0
        width += width & 1
    if (trg not in ("cubemap", "biatan6") and height and width > 4 and height > 4 and "single" not in kw
            and rng.random() < .12):  # cropped output: p-line with an S clause (no camera rotation on that route)
        x0, y0 = int(rng.integers(0, width // 2)), int(rng.integers(0, height // 2))
        kw.update(crop_out=(x0, int(rng.integers(x0 + 1, width + 1)), y0, int(rng.integers(y0 + 1, height + 1))),
                  yaw=0.0, pitch=0.0, roll=0.0)
    if "solo" not in kw and "single" not in kw and rng.random() < .12:  # --mask_for, half of them reduced by --nchannels
        kw["mask_for"] = int(rng.integers(0, nf))
        if rng.random() < .5:
            kw["out_channels"] = int(rng.choice([1, 2]))
    if "mask_for" not in kw and "single" not in kw and rng.random() < .08:  # --nchannels on an ordinary job: repix_t both ways
        kw["out_channels"] = int(rng.integers(1, 5))
    return Job(facets, trg, hfov, width, height, **kw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    bad = crashed = ok = skipped = known = 0
    for k in range(a.n):
        job = random_job(rng)
        desc = "%s<-%s d%d tw%d %dx%d %s" % (job.projection, "+".join("%s%dx%d" % ((f.projection,) + f.native_shape()[:2])
                                                                     for f in job.facets), job.degree,
                                              job.twine, job.width, job.height,
                                              " ".join(x for x in ("hdr" if job.synopsis != "panorama" else "",
                                                                   "solo%d" % job.solo if job.solo >= 0 else "",
                                                                   "single%d" % job.single if job.single >= 0 else "") if x))
        pm = None
        for attempt in range(4):
            try:
                pm = harness.reference_render(job, "pm")
                break
            except RuntimeError as e:
                if "(-11)" in str(e) or "(-6)" in str(e):
                    continue
                pm = None
                break
            except Exception:
                break
        if pm is None:
            skipped += 1
            continue
        try:
            orc = harness.oracle_render(job)
        except Exception as e:
            crashed += 1
            print("#%d ORACLE ERROR %s: %s" % (k, desc, str(e)[:120]), flush=True)
            continue
        if pm.shape != orc.shape or not np.array_equal(pm, orc, equal_nan=True):
            # the two categories DESIGN.md lists as not pinned / refused by the library
            odd_cube = any(f.projection in ("cubemap", "biatan6") and f.native_shape()[0] % 2 for f in job.facets)
            # a full-sphere lat/lon image lower than its over-the-pole brace (environment.h:473-516 folds twice)
            tiny = any(f.projection == "spherical" and f.hfov == 360.0 and f.native_shape()[0] == 2 * f.native_shape()[1]
                       and f.native_shape()[1] < job.degree // 2 + 1 for f in job.facets)
            thin = False  # cubemap support frame narrower than the spline window: undefined in the reference
            for f in job.facets:
                if f.projection in ("cubemap", "biatan6"):
                    import ctypes as C
                    from envutil_b200 import capi
                    oi, od = (C.c_int32 * 4)(), (C.c_double * 4)()
                    capi.load().eu_cubemap_metrics(f.native_shape()[0], f.hfov * 3.141592653589793 / 180.0, job.support_min,
                                                   job.tile_size, oi, od)
                    frame_l = oi[1]
                    frame_r = oi[2]
                    import math
                    inherent = int(math.trunc(od[1] * (math.tan(math.radians(f.hfov) / 2.0) - 1.0))) if f.hfov > 90.0 else 0
                    thin = thin or min(frame_l, frame_r) + inherent < job.degree // 2 + 1
            if odd_cube or tiny or thin:
                known += 1
                continue
            bad += 1
            nd = int((pm != orc).sum()) if pm.shape == orc.shape else -1
            print("#%d DIFF %s: %d of %d values, max abs %.3g" % (k, desc, nd, pm.size,
                  float(np.nanmax(np.abs(pm - orc))) if nd > 0 else 0.0), flush=True)
        else:
            ok += 1
    print("jobs %d: identical %d, different %d (+ %d in the known categories: odd cube face width, full-sphere "
          "image lower than its pole brace, cubemap support frame narrower than the spline window), oracle errors %d, reference refused/crashed %d" % (a.n, ok, bad, known, crashed, skipped))


if __name__ == "__main__":
    main()
