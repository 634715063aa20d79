"""Measurement tooling that runs on the CPU: the exact algorithmic-byte count."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_count_touched_small_c1_and_c2():
    """tools/count_touched.py counts the distinct container texels a job's spline windows read, from the
    oracle's tap addresses. At 1/8 size: a 90-degree rectilinear view touches about 7.5 % of a lat/lon
    source (7.48 %), a full sphere touches every face texel of a cubemap and a little of
    its support frame."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "count_touched.py"), "--configs", "C1,C2", "--scale", "8"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = {d["config"]: d for d in (json.loads(l) for l in r.stdout.splitlines() if l.startswith("{"))}
    c1, c2 = rows["C1"], rows["C2"]
    assert 0.070 < c1["touched_fraction"] < 0.080
    assert c1["algorithmic_bytes_exact"] == c1["out_px"] * 12 + c1["touched_texels"] * 12
    face = 2048 // 8
    assert 6 * face * face <= c2["touched_texels"] <= c2["container_texels"]


def test_exact_counts_are_what_the_workloads_use():
    sys.path.insert(0, ROOT)
    from envutil_b200 import workloads
    assert workloads.c1(1)[1] == 1920 * 1080 * 12 + workloads.EXACT_TOUCHED["C1"] * 12
    assert workloads.c2(1)[1] == 8192 * 4096 * 12 + workloads.EXACT_TOUCHED["C2"] * 12 == 705546384
