#!/usr/bin/env python3
"""Whole-frame parity of the CUDA path at BASELINE.json's FULL sizes against the unmodified reference
binaries under oracle/_ref/ (run here, on the box's host cores) - every config, every pixel.

  python tools/full_parity.py [--configs C1,C2,C3a,C3b,C4,C5A,C5B] [--out gpurun_out/parity_full.json]
writes <out>_exact.json and <out>_contracted.json: the same frames rendered with both arithmetics of the library
(EU_OPT_CONTRACTED), each compared with the same reference frames.

Per config one record:
  vs_pinned   GPU frame against oracle/_ref/envutil_ref_pm (the reference sources with the elementary
              functions of include/eu_math.h interposed for libm's): max / RMS of |gpu-ref| / max(|ref|, 1e-3)
              and the number of floats that differ at all (the contract of the default arithmetic is 0)
  vs_libm     the same frame against oracle/_ref/envutil_ref (stock libm): max / RMS, the count of floats
              beyond 1e-5, and the same figures with the tie band masked (pixels whose cube-face
              or winning-facet choice is within 8 ulp of flipping - eu_debug_tie_plane - and their count)
  ref_self    envutil_ref_pm against envutil_ref: how far two builds of the REFERENCE are apart on this job
The reference is test infrastructure; nothing here is used by the product path. /root/reference is not read.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import harness  # noqa: E402
from envutil_b200 import euf, workloads  # noqa: E402

EPS = 1e-3
TOL = 1e-5


def rel_stats(a, b, mask=None, chunk_rows=256):
    """max / RMS of |a-b| / max(|b|, EPS) over all floats (row chunks on a few threads: the frames are GB-sized),
    count of floats that differ, count beyond TOL. mask: H x W bool, pixels to leave out."""
    from concurrent.futures import ThreadPoolExecutor

    def part(y0):
        x = a[y0:y0 + chunk_rows].astype(np.float64)
        y = b[y0:y0 + chunk_rows].astype(np.float64)
        d = np.abs(x - y)
        rel = d / np.maximum(np.abs(y), EPS)
        if mask is not None:
            keep = ~mask[y0:y0 + chunk_rows]
            rel = rel[keep]
            d = d[keep]
        if rel.size == 0:
            return 0.0, 0.0, 0.0, 0, 0, 0
        return (float(rel.max()), float(d.max()), float((rel * rel).sum()), int(rel.size), int((d != 0).sum()),
                int((rel > TOL).sum()))
    with ThreadPoolExecutor(max_workers=max(1, min(16, os.cpu_count() or 1))) as ex:
        parts = list(ex.map(part, range(0, a.shape[0], chunk_rows)))
    n = sum(p[3] for p in parts)
    return {"max_rel": max(p[0] for p in parts), "rms_rel": (sum(p[2] for p in parts) / max(n, 1)) ** 0.5,
            "max_abs": max(p[1] for p in parts), "n_diff": sum(p[4] for p in parts),
            "n_beyond_1e-5": sum(p[5] for p in parts), "n": n}


def run_reference(job, kind, workdir, paths=None, tag="out"):
    """One run of an unmodified reference binary on the job; returns (frame, seconds)."""
    exe = harness.ref_binary(kind)
    if not exe:
        raise RuntimeError("oracle/_ref/%s is missing" % kind)
    if paths is None:
        paths = []
        for i, f in enumerate(job.facets):
            p = os.path.join(workdir, "facet%d.euf" % i)
            euf.write_euf(p, f.image)
            paths.append(p)
    outp = os.path.join(workdir, "%s_%s.euf" % (tag, kind))
    t0 = time.perf_counter()
    r = subprocess.run([exe] + job.cli_args(paths, outp), capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0 or not os.path.exists(outp):
        raise RuntimeError("reference (%s) failed: %s\n%s" % (kind, r.stdout[-1500:], r.stderr[-1500:]))
    img = euf.read_euf(outp)
    os.unlink(outp)
    return img, dt, paths


ARITHMETICS = ("exact", "contracted")


def parity_of(engine, job, workdir, want_libm=True, keep=False, log=None):
    """GPU frames of the job - both arithmetics of the library - against both reference builds.
    Returns ({arithmetic: record}, pinned reference frame or None)."""
    st = job.structs(engine.lib)
    hs = engine.stage(job, st)
    gpu = {}
    try:
        for ar in ARITHMETICS:
            job.contracted = ar == "contracted"
            t0 = time.perf_counter()
            gpu[ar] = (engine.render(job, sources=hs, structs=job.structs(engine.lib)), time.perf_counter() - t0)
        job.contracted = None
        tie = engine.tie_plane(job, hs, st, 8)
    finally:
        job.contracted = None
        engine.release(hs)
    pm, pm_s, paths = run_reference(job, "pm", workdir)
    lm, lm_s = (None, None)
    if want_libm:
        lm, lm_s, _ = run_reference(job, "libm", workdir, paths)
    for p in paths:
        os.unlink(p)
    ref_self = rel_stats(pm, lm) if want_libm else None
    recs = {}
    for ar in ARITHMETICS:
        frame, gpu_s = gpu[ar]
        rec = {"config": job.name, "out": "%dx%d" % (frame.shape[1], frame.shape[0]), "floats": int(frame.size),
               "vs_pinned": rel_stats(frame, pm), "ref_pm_s": pm_s, "gpu_render_and_download_s": gpu_s,
               "tie_pixels": int(tie.sum()) if tie is not None else 0}
        if want_libm:
            rec["vs_libm"] = rel_stats(frame, lm)
            if tie is not None and tie.any():
                m = rel_stats(frame, lm, mask=tie.astype(bool))
                rec["vs_libm"]["masked"] = {"pixels": int(tie.sum()), "max_rel": m["max_rel"], "rms_rel": m["rms_rel"],
                                            "n_beyond_1e-5": m["n_beyond_1e-5"]}
            else:
                rec["vs_libm"]["masked"] = {"pixels": 0, "max_rel": rec["vs_libm"]["max_rel"],
                                            "rms_rel": rec["vs_libm"]["rms_rel"],
                                            "n_beyond_1e-5": rec["vs_libm"]["n_beyond_1e-5"]}
            rec["ref_self"] = ref_self
            rec["ref_libm_s"] = lm_s
        recs[ar] = rec
        if log:
            log("%s %s vs_pinned %s" % (job.name, ar, rec["vs_pinned"]))
            if want_libm:
                log("%s %s vs_libm %s" % (job.name, ar, rec["vs_libm"]))
    back = {ar: gpu[ar][0] for ar in ARITHMETICS} if keep else None
    return recs, (pm if keep else None), back


def configs_iter(engine, want, workdir, scale=1, want_libm=True, log=None):
    """Yields {arithmetic: record} per config, in BASELINE order. C3b takes the reference's own C3a output as its
    input (both legs of the round trip are then compared on identical inputs), C5B the reference's stage-A results."""
    if "C1" in want:
        job, _ = workloads.c1(scale)
        yield parity_of(engine, job, workdir, want_libm, log=log)[0]
    if "C2" in want:
        job, _ = workloads.c2(scale)
        yield parity_of(engine, job, workdir, want_libm, log=log)[0]
    if "C3a" in want or "C3b" in want:
        job, _ = workloads.c3a(scale)
        recs, pm, _ = parity_of(engine, job, workdir, want_libm, keep=True, log=log)
        ll = job.facets[0].image
        if "C3a" in want:
            yield recs
        if "C3b" in want:
            job2, _ = workloads.c3b(pm)
            recs2, back_ref, back = parity_of(engine, job2, workdir, want_libm, keep=True, log=log)
            e_ref = rel_round_trip(back_ref, ll)
            for ar in ARITHMETICS:
                recs2[ar]["round_trip"] = {"gpu": rel_round_trip(back[ar], ll), "reference": e_ref}
            yield recs2
    if "C4" in want:
        job, _ = workloads.c4(scale)
        yield parity_of(engine, job, workdir, want_libm, log=log)[0]
    if "C5A" in want or "C5B" in want:
        fs = workloads.c5_facets(scale)
        merged, yaws, per = [], [], []
        for k in range(0, len(fs), 3):
            job, _ = workloads.c5_stage_a(fs[k:k + 3])
            job.name = "C5A[%d]" % (k // 3)
            # every position for the reference's merged image (stage B's input); libm only on the first
            recs, pm, _ = parity_of(engine, job, workdir, want_libm and k == 0, keep=True, log=log)
            per.append(recs)
            merged.append(pm)
            yaws.append(fs[k].yaw)
        if "C5A" in want:
            out = {}
            for ar in ARITHMETICS:
                rs = [r[ar] for r in per]
                agg = {"config": "C5A", "out": "6 x %s" % rs[0]["out"], "floats": sum(r["floats"] for r in rs),
                       "vs_pinned": {"max_rel": max(r["vs_pinned"]["max_rel"] for r in rs),
                                     "rms_rel": float(np.sqrt(np.mean([r["vs_pinned"]["rms_rel"] ** 2 for r in rs]))),
                                     "max_abs": max(r["vs_pinned"]["max_abs"] for r in rs),
                                     "n_diff": sum(r["vs_pinned"]["n_diff"] for r in rs),
                                     "n_beyond_1e-5": sum(r["vs_pinned"]["n_beyond_1e-5"] for r in rs),
                                     "n": sum(r["vs_pinned"]["n"] for r in rs)},
                       "tie_pixels": 0, "positions": rs}
                if "vs_libm" in rs[0]:
                    agg["vs_libm"] = dict(rs[0]["vs_libm"], note="position 0 only")
                    agg["ref_self"] = rs[0]["ref_self"]
                out[ar] = agg
            yield out
        if "C5B" in want:
            del fs
            job, _ = workloads.c5_stage_b(merged, yaws, scale=scale)
            yield parity_of(engine, job, workdir, want_libm, log=log)[0]


def rel_round_trip(back, original):
    h = back.shape[0]
    mx, ss = 0.0, 0.0
    for y0 in range(0, h, 512):
        d = np.abs(back[y0:y0 + 512].astype(np.float64) - original[y0:y0 + 512])
        mx = max(mx, float(d.max()))
        ss += float((d * d).sum())
    return {"max_abs": mx, "rms": (ss / back.size) ** 0.5}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3a,C3b,C4,C5A,C5B")
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_full.json"))
    ap.add_argument("--no-libm", action="store_true")
    a = ap.parse_args()
    from envutil_b200.engine import Engine
    eng = Engine(0)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    workdir = tempfile.mkdtemp(prefix="euparity_", dir=base)
    recs = {ar: [] for ar in ARITHMETICS}
    t0 = time.time()

    def log(s):
        print("[%6.1f s] %s" % (time.time() - t0, s), file=sys.stderr, flush=True)
    try:
        for both in configs_iter(eng, a.configs.split(","), workdir, a.scale, not a.no_libm, log):
            for ar in ARITHMETICS:
                recs[ar].append(both[ar])
            print(json.dumps(both), flush=True)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
        eng.close()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    stem = a.out[:-5] if a.out.endswith(".json") else a.out
    for ar in ARITHMETICS:
        with open("%s_%s.json" % (stem, ar), "w") as f:
            json.dump({"arithmetic": ar, "scale": a.scale, "eps": EPS, "tolerance": TOL, "cores": os.cpu_count(),
                       "configs": recs[ar], "seconds": time.time() - t0,
                       "reference_builds": {"pinned": "oracle/_ref/envutil_ref_pm (-O2 -ffp-contract=off, eu_math.h interposed)",
                                            "libm": "oracle/_ref/envutil_ref (-O2 -ffp-contract=off, stock libm)"}}, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
