#!/usr/bin/env python3
"""Random jobs through `envutil_b200_cli --dry_run` (the C++ host: option table, PTO lines, Eev -> brighten,
twining set-up) against the Python marshalling the parity tests use (envutil_b200/job.py, which is pinned
to the reference through the golden outputs): target, extent, every facet's derived geometry and the twining
taps must agree to the last bit. CPU only.

  python tools/fuzz_cli_dry_run.py [--n 300] [--seed 5]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

from envutil_b200 import euf  # noqa: E402
from fuzz_oracle_vs_reference import random_job  # noqa: E402

CLI = os.path.join(ROOT, "envutil_b200", "envutil_b200_cli")


def floats(line, key):
    m = re.search(key + r" ((?:[-+0-9.eEinfa]+ ?)+)", line)
    return [float(v) for v in m.group(1).split()]


def check(job, d):
    paths = []
    for i, f in enumerate(job.facets):
        paths.append(os.path.join(d, "facet%d.euf" % i))
        euf.write_euf(paths[-1], f.image)
    try:
        args = job.cli_args(paths, os.path.join(d, "out.euf"))
    except KeyError:  # PTO i-lines cannot name a cubemap: such a mix has no command line
        return "refused"
    r = subprocess.run([CLI] + args + ["--dry_run"], capture_output=True, text=True)
    try:
        t, fa, o, taps, ntaps = job.structs()
        ok_py = True
    except Exception:
        ok_py = False
    if r.returncode != 0 or not ok_py:
        return "refused" if (r.returncode != 0) == (not ok_py) else "one side refused: cli rc %d, python %s: %s" % (
            r.returncode, ok_py, r.stderr.strip()[:120])
    lines = r.stdout.splitlines()
    tl = [l for l in lines if l.startswith("target ")][0]
    if "%dx%d nch %d " % (t.width, t.height, t.nchannels) not in tl:
        return "target size / channels: " + tl
    if [floats(tl, k)[0] for k in ("hfov", "yaw", "pitch", "roll")] != [t.hfov, t.yaw, t.pitch, t.roll]:
        return "target angles: " + tl
    el = [l for l in lines if l.startswith("extent ")][0]
    if floats(el, "extent") != [t.x0, t.x1, t.y0, t.y1] or floats(el, "step")[0] != t.step:
        return "extent: " + el
    fl = [l for l in lines if l.startswith("facet ")]
    if len(fl) != len(job.facets):
        return "facet count"
    for i, l in enumerate(fl):
        if (floats(l, "hfov")[0] != fa[i].hfov or floats(l, "ypr") != [fa[i].yaw, fa[i].pitch, fa[i].roll]
                or floats(l, " step")[0] != fa[i].step or np.float32(floats(l, "brighten")[0]) != np.float32(fa[i].brighten)
                or floats(l, "shift") != [fa[i].shift_h, fa[i].shift_v] or floats(l, "shear") != [fa[i].shear_g, fa[i].shear_t]
                or int(l.rsplit("masked ", 1)[1]) != fa[i].masked):
            return "facet %d: %s" % (i, l[:200])
    tp = [l for l in lines if l.startswith("tap ")]
    if len(tp) != ntaps:
        return "tap count %d vs %d" % (len(tp), ntaps)
    for k, l in enumerate(tp):
        x, y, w = (np.float32(v) for v in l.split()[1:])
        if (x, y, w) != (np.float32(taps[k].x), np.float32(taps[k].y), np.float32(taps[k].w)):
            return "tap %d" % k
    dl = [l for l in lines if l.startswith("degree ")][0]
    if "solo %d" % o.solo not in dl or "degree %d " % o.spline_degree not in dl:
        return "options: " + dl
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--seed", type=int, default=5)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    same = bad = refused = 0
    for k in range(a.n):
        job = random_job(rng)
        with tempfile.TemporaryDirectory(prefix="eucli_") as d:
            res = check(job, d)
        if res is None:
            same += 1
        elif res == "refused":
            refused += 1
        else:
            bad += 1
            print("#%d %s<-%s d%d tw%d: %s" % (k, job.projection, "+".join(f.projection for f in job.facets), job.degree,
                                              job.twine, res), flush=True)
    print("jobs %d: identical marshalling %d, different %d, refused by both %d" % (a.n, same, bad, refused))


if __name__ == "__main__":
    main()
