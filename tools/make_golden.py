#!/usr/bin/env python3
"""Generate tests/golden/: reference outputs for the small parity jobs of tests/jobs.py.

Runs every job through the UNMODIFIED reference built by oracle/Makefile:
  oracle/_ref/envutil_ref_pm   reference sources + the elementary functions of include/eu_math.h
                               (the numerical contract of the B200 back-end)  -> golden outputs
  oracle/_ref/envutil_ref      reference sources + glibc libm                 -> drift statistics
and writes
  tests/golden/manifest.json   per job: shape, sha256 of the float32 output bytes of the
                               pinned-math build, and max/RMS relative difference between the
                               two builds (how far two legitimate builds of the reference are
                               apart - the context for the 1e-5 tolerance)
  tests/golden/<job>.npz       full pinned-math output for the jobs listed in FULL (small)
Only this container can run it (it needs oracle/_ref, which is built from /root/reference);
tests read the committed files.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
import jobs  # noqa: E402

FULL = ["ll_rect_d1", "ll_rect_d3_rot", "cm_sph_d3", "ba6_sph_d1", "ll_ba6_d1", "ll_fish_d1_tw4",
        "voronoi4_sph_d1", "hdr3_rect_d1", "lens3_voronoi_sph_d1", "ll_cyl_d1_tw2"]


def _retry(job, kind, n=5):
    """The reference evaluates masked-out lanes at an uninitialised coordinate before zeroing them
    (environment.h:592,1190-1193); with NaN rays (translation) that occasionally reads outside
    its arrays and crashes. The result of a run that completes does not depend on it: retry."""
    for k in range(n):
        try:
            return harness.reference_render(job, kind)
        except RuntimeError as e:
            if k == n - 1 or "(-11)" not in str(e):
                raise


def main():
    os.makedirs(harness.GOLDEN, exist_ok=True)
    manifest = {}
    all_jobs = dict(jobs.JOBS, **jobs.EDGE_JOBS)
    for name in sorted(all_jobs):
        job = all_jobs[name]
        pm = _retry(job, "pm")
        lm = _retry(job, "libm")
        c = harness.compare(lm, pm)
        manifest[name] = {
            "shape": list(pm.shape),
            "sha256": hashlib.sha256(np.ascontiguousarray(pm, dtype="<f4").tobytes()).hexdigest(),
            "libm_vs_pinned": {k: c[k] for k in ("max_rel", "rms_rel", "max_abs", "n_diff")},
        }
        if name in FULL:
            np.savez_compressed(os.path.join(harness.GOLDEN, name + ".npz"), out=pm)
        print(name, manifest[name]["sha256"][:12], c["max_rel"])
    import tempfile
    for name, (base, twf, extra) in sorted(jobs.CLI_EXTRAS.items()):
        with tempfile.TemporaryDirectory() as d:
            tp = os.path.join(d, "filter.twf")
            open(tp, "w").write(twf)
            pm = harness.reference_render(jobs.JOBS[base], "pm", extra_args=["--twf_file", tp] + extra)
        manifest[name] = {"shape": list(pm.shape),
                          "sha256": hashlib.sha256(np.ascontiguousarray(pm, dtype="<f4").tobytes()).hexdigest(),
                          "cli_extra": True}
        print(name, manifest[name]["sha256"][:12])
    import subprocess
    from envutil_b200 import euf
    for name, (base, extra) in sorted(jobs.SPLITS.items()):  # --split: one output per facet
        job = jobs.JOBS[base]
        for attempt in range(6):
            with tempfile.TemporaryDirectory(prefix="euref_") as d:
                paths = []
                for i, f in enumerate(job.facets):
                    paths.append(os.path.join(d, "facet%d.euf" % i))
                    euf.write_euf(paths[-1], f.image)
                args = job.cli_args(paths, "unused.euf")
                k = args.index("--output")
                del args[k:k + 2]
                r = subprocess.run([harness.ref_binary("pm")] + args + extra + ["--split", os.path.join(d, "re%02d.euf")],
                                   capture_output=True, text=True)
                if r.returncode == -11 and attempt < 5:  # see _retry
                    continue
                assert r.returncode == 0, r.stderr[-2000:]
                outs = {}
                for i in range(len(job.facets)):
                    q = os.path.join(d, "re%02d.euf" % i)
                    if os.path.exists(q):
                        img = euf.read_euf(q)
                        outs[str(i)] = {"shape": list(img.shape),
                                        "sha256": hashlib.sha256(np.ascontiguousarray(img, dtype="<f4").tobytes()).hexdigest()}
                manifest[name] = {"split": True, "base": base, "extra": extra, "outputs": outs}
                print(name, sorted(outs))
                break
    with open(os.path.join(harness.GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
