#!/usr/bin/env python3
"""Re-run ONE job of the random sweep (tools/fuzz_gpu_vs_oracle.py) in isolation and in variants, to localise a
difference between the kernels and the oracle: python tools/debug_fuzz_job.py --seed 12 --index 136"""
import argparse
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]

import numpy as np  # noqa: E402

import harness  # noqa: E402
from envutil_b200.engine import Engine  # noqa: E402
from fuzz_oracle_vs_reference import random_job  # noqa: E402


def report(eng, job, label):
    try:
        out = eng.render(job)
    except RuntimeError as e:
        print("%-34s refused: %s" % (label, str(e)[:100]), flush=True)
        return
    ref = harness.oracle_render(job)
    d = out != ref
    c = harness.compare(out, ref)
    msg = "%-34s n_diff %6d of %6d  max_rel %.3g max_abs %.3g" % (label, c["n_diff"], c["n"], c["max_rel"], c["max_abs"])
    if d.any():
        ys, xs, cs = np.nonzero(d)
        msg += "  rows %d..%d cols %d..%d first (%d,%d,%d): gpu %r oracle %r" % (
            ys.min(), ys.max(), xs.min(), xs.max(), ys[0], xs[0], cs[0], float(out[ys[0], xs[0], cs[0]]),
            float(ref[ys[0], xs[0], cs[0]]))
    print(msg, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=12)
    ap.add_argument("--index", type=int, default=136)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    for _ in range(a.index + 1):
        job = random_job(rng)
    eng = Engine(0)
    report(eng, job, "as drawn")
    report(eng, job, "as drawn, again")
    for name, kw in (("no_spec", dict(no_spec=True)), ("no_tiles", dict(no_tiles=True)), ("twine 0", dict(twine=0)),
                     ("degree 1", dict(degree=1)), ("single -1", dict(single=-1)),
                     ("twine 0, single -1", dict(twine=0, single=-1))):
        j = copy.copy(job)
        for k, v in kw.items():
            setattr(j, k, v)
        report(eng, j, name)
    for k in range(len(job.facets)):
        j = copy.copy(job)
        j.solo = k
        report(eng, j, "solo %d" % k)
        j = copy.copy(job)
        j.solo, j.twine = k, 0
        report(eng, j, "solo %d, twine 0" % k)
    for k in range(len(job.facets)):
        j = copy.copy(job)
        j.facets = [f for i, f in enumerate(job.facets) if i != k]
        if job.single >= 0:
            if k == job.single:
                continue
            j.single = job.single - (1 if k < job.single else 0)
        report(eng, j, "without facet %d" % k)
    eng.close()


if __name__ == "__main__":
    main()
