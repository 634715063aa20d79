#!/usr/bin/env python3
"""The same random jobs as tools/fuzz_oracle_vs_reference.py, through the CUDA path and the oracle, compared
bit for bit (runs on a B200 box; the oracle is the checker). Jobs the library refuses (EU_ERR_UNSUPPORTED /
EU_ERR_ARGUMENT) are counted, not compared. First run on a B200 at the start of round 2 (profiles/r02a_parity_summary.txt).

  python tools/fuzz_gpu_vs_oracle.py [--n 500] [--seed 11]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

import harness  # noqa: E402
from envutil_b200.engine import Engine  # noqa: E402
from fuzz_oracle_vs_reference import random_job  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=500)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--contracted", type=int, default=0,
                    help="1: the contracted arithmetic (EU_OPT_CONTRACTED) against the oracle's restatement of it")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    eng = Engine(0)
    same = diff = refused = known = 0
    for k in range(a.n):
        job = random_job(rng)
        job.contracted = bool(a.contracted)
        desc = "%s<-%s d%d tw%d %dx%d" % (job.projection, "+".join("%s%dx%d" % ((f.projection,) + f.native_shape()[:2])
                                                                    for f in job.facets), job.degree, job.twine,
                                          job.width, job.height)
        try:
            out = eng.render(job)
        except RuntimeError as e:
            refused += 1
            if "status -2" not in str(e) and "status -1" not in str(e):
                print("#%d LIBRARY ERROR %s: %s" % (k, desc, str(e)[:160]), flush=True)
            continue
        try:
            ref = harness.oracle_render(job, contracted=bool(a.contracted))
        except Exception as e:
            print("#%d ORACLE ERROR %s: %s" % (k, desc, str(e)[:120]), flush=True)
            continue
        if out.shape == ref.shape and np.array_equal(out, ref, equal_nan=True):
            same += 1
            continue
        odd_cube = any(f.projection in ("cubemap", "biatan6") and f.native_shape()[0] % 2 for f in job.facets)
        if odd_cube:  # the support fill of odd cube faces is order-dependent
            known += 1
            continue
        diff += 1
        nd = int((out != ref).sum()) if out.shape == ref.shape else -1
        print("#%d DIFF %s: %d of %d values" % (k, desc, nd, ref.size), flush=True)
    eng.close()
    print("jobs %d: identical %d, different %d (+ %d with odd cube faces), refused by the library %d"
          % (a.n, same, diff, known, refused))


if __name__ == "__main__":
    main()
