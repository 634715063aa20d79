#!/usr/bin/env python3
"""Whole-frame parity of the CUDA path at BASELINE.json's FULL sizes against the unmodified reference
binaries under oracle/_ref/ (run here, on the box's host cores) - every config, every pixel.

  python tools/full_parity.py [--configs C1,C2,C3a,C3b,C4,C5A,C5B] [--out gpurun_out/parity_full.json]

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


def rel_stats(a, b, mask=None, chunk_rows=512):
    """max / RMS of |a-b| / max(|b|, EPS) over all floats (row chunks: the frames are GB-sized), count of
    floats that differ, count beyond TOL. mask: H x W bool, pixels to leave out."""
    h = a.shape[0]
    mx, ss, n, ndiff, nbeyond, mabs = 0.0, 0.0, 0, 0, 0, 0.0
    for y0 in range(0, h, chunk_rows):
        x = a[y0:y0 + chunk_rows].astype(np.float64)
        y = b[y0:y0 + chunk_rows].astype(np.float64)
        d = np.abs(x - y)
        rel = d / np.maximum(np.abs(y), EPS)
        if mask is not None:
            keep = ~mask[y0:y0 + chunk_rows]
            rel = rel[keep]
            d = d[keep]
        if rel.size == 0:
            continue
        mx = max(mx, float(rel.max()))
        mabs = max(mabs, float(d.max()))
        ss += float((rel * rel).sum())
        n += rel.size
        ndiff += int((d != 0).sum())
        nbeyond += int((rel > TOL).sum())
    return {"max_rel": mx, "rms_rel": (ss / max(n, 1)) ** 0.5, "max_abs": mabs, "n_diff": ndiff, "n_beyond_1e-5": nbeyond,
            "n": n}


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


def parity_of(engine, job, workdir, want_libm=True, keep_gpu=False, log=None):
    """GPU frame of the job against both reference builds. Returns the record (and the GPU frame)."""
    st = job.structs(engine.lib)
    t0 = time.perf_counter()
    hs = engine.stage(job, st)
    try:
        gpu = engine.render(job, sources=hs, structs=st)
        tie = engine.tie_plane(job, hs, st, 8)
    finally:
        engine.release(hs)
    gpu_s = time.perf_counter() - t0
    pm, pm_s, paths = run_reference(job, "pm", workdir)
    rec = {"config": job.name, "out": "%dx%d" % (gpu.shape[1], gpu.shape[0]), "floats": int(gpu.size),
           "vs_pinned": rel_stats(gpu, pm), "ref_pm_s": pm_s, "gpu_path_s": gpu_s,
           "tie_pixels": int(tie.sum()) if tie is not None else 0}
    if log:
        log("%s vs_pinned %s" % (job.name, rec["vs_pinned"]))
    if want_libm:
        lm, lm_s, _ = run_reference(job, "libm", workdir, paths)
        rec["vs_libm"] = rel_stats(gpu, lm)
        if tie is not None and tie.any():
            m = rel_stats(gpu, lm, mask=tie.astype(bool))
            rec["vs_libm"]["masked"] = {"pixels": int(tie.sum()), "max_rel": m["max_rel"], "rms_rel": m["rms_rel"],
                                        "n_beyond_1e-5": m["n_beyond_1e-5"]}
        else:
            rec["vs_libm"]["masked"] = {"pixels": 0, "max_rel": rec["vs_libm"]["max_rel"],
                                        "rms_rel": rec["vs_libm"]["rms_rel"],
                                        "n_beyond_1e-5": rec["vs_libm"]["n_beyond_1e-5"]}
        rec["ref_self"] = rel_stats(pm, lm)
        rec["ref_libm_s"] = lm_s
        if log:
            log("%s vs_libm %s" % (job.name, rec["vs_libm"]))
        del lm
    for p in paths:
        os.unlink(p)
    return (rec, gpu, pm) if keep_gpu else (rec, None, None)


def configs_iter(engine, want, workdir, scale=1, want_libm=True, log=None):
    """Yields one record per config, in BASELINE order. C3b takes the reference's own C3a output as its
    input (both legs of the round trip are then compared on identical inputs), C5B the reference's
    stage-A results."""
    if "C1" in want:
        job, _ = workloads.c1(scale)
        yield parity_of(engine, job, workdir, want_libm, log=log)[0]
    if "C2" in want:
        job, _ = workloads.c2(scale)
        yield parity_of(engine, job, workdir, want_libm, log=log)[0]
    if "C3a" in want or "C3b" in want:
        job, _ = workloads.c3a(scale)
        rec, gpu, pm = parity_of(engine, job, workdir, want_libm, keep_gpu=True, log=log)
        ll = job.facets[0].image
        if "C3a" in want:
            yield rec
        if "C3b" in want:
            del gpu
            job2, _ = workloads.c3b(pm)
            rec2, back, back_ref = parity_of(engine, job2, workdir, want_libm, keep_gpu=True, log=log)
            e = rel_round_trip(back, ll)
            e_ref = rel_round_trip(back_ref, ll)
            rec2["round_trip"] = {"gpu": e, "reference": e_ref}
            yield rec2
    if "C4" in want:
        job, _ = workloads.c4(scale)
        yield parity_of(engine, job, workdir, want_libm, log=log)[0]
    if "C5A" in want or "C5B" in want:
        fs = workloads.c5_facets(scale)
        merged, yaws, recs = [], [], []
        for k in range(0, len(fs), 3):
            job, _ = workloads.c5_stage_a(fs[k:k + 3])
            job.name = "C5A[%d]" % (k // 3)
            # every position for the reference's merged image (stage B's input); libm only on the first
            rec, gpu, pm = parity_of(engine, job, workdir, want_libm and k == 0, keep_gpu=True, log=log)
            recs.append(rec)
            merged.append(pm)
            yaws.append(fs[k].yaw)
            del gpu
        if "C5A" in want:
            agg = {"config": "C5A", "out": "6 x %s" % recs[0]["out"], "floats": sum(r["floats"] for r in recs),
                   "vs_pinned": {"max_rel": max(r["vs_pinned"]["max_rel"] for r in recs),
                                 "rms_rel": float(np.sqrt(np.mean([r["vs_pinned"]["rms_rel"] ** 2 for r in recs]))),
                                 "n_diff": sum(r["vs_pinned"]["n_diff"] for r in recs),
                                 "n_beyond_1e-5": sum(r["vs_pinned"]["n_beyond_1e-5"] for r in recs),
                                 "n": sum(r["vs_pinned"]["n"] for r in recs)},
                   "tie_pixels": 0, "positions": recs}
            if "vs_libm" in recs[0]:
                agg["vs_libm"] = dict(recs[0]["vs_libm"], note="position 0 only")
                agg["ref_self"] = recs[0]["ref_self"]
            yield agg
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
    from envutil_b200 import capi
    from envutil_b200.engine import Engine
    eng = Engine(0)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    workdir = tempfile.mkdtemp(prefix="euparity_", dir=base)
    recs = []
    t0 = time.time()

    def log(s):
        print("[%6.1f s] %s" % (time.time() - t0, s), file=sys.stderr, flush=True)
    try:
        for rec in configs_iter(eng, a.configs.split(","), workdir, a.scale, not a.no_libm, log):
            recs.append(rec)
            print(json.dumps(rec), flush=True)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
        eng.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"arithmetic": capi.ARITHMETIC if hasattr(capi, "ARITHMETIC") else "exact", "scale": a.scale,
                   "eps": EPS, "tolerance": TOL, "cores": os.cpu_count(), "configs": recs,
                   "reference_builds": {"pinned": "oracle/_ref/envutil_ref_pm (-O2 -ffp-contract=off, eu_math.h interposed)",
                                        "libm": "oracle/_ref/envutil_ref (-O2 -ffp-contract=off, stock libm)"}}, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
