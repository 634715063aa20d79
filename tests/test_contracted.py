"""The opt-in arithmetic variant (libenvutil_b200_fma.so, EU_ARITHMETIC=contracted; include/envutil_b200.h
eu_render_arithmetic): fused multiply-adds in the b-spline window evaluation and the twining accumulation,
nothing else. The default library stays bit-identical to the reference's parity build; this variant is what
a reference compiled with g++'s default -ffp-contract=fast on FMA hardware resembles, and it has to stay
within BASELINE.json's tolerance (1e-5 relative) of the pinned reference.

CPU: the library variant loads and names its arithmetic; the oracle's restatement of the variant
(orc_set_arithmetic) keeps every index plane and differs from the pinned arithmetic by at most 2.6e-6
relative (hdr_merge's near-zero denominators: 1.6e-5), RMS <= 2.7e-7, over all 98 small jobs.
GPU (opt-in, no GPU run has seen it yet): the contracted kernels equal the contracted oracle bit for bit.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import harness
import jobs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# a cross-section that runs in seconds: bilinear, quadratic, cubic, quintic; cubemaps, twining, voronoi,
# alpha compositing, translated facets (run-time evaluator), cropped output
SUBSET = ["ll_rect_d1_rot", "ll_rect_d2", "ll_rect_d3_rot", "ll_rect_d5", "cm_sph_d3", "ba6_sph_d1", "ll_fish_d1_tw4",
          "voronoi4_sph_d3_rot", "rgba4_voronoi_sph_d3", "tr1_sph_d1", "crop1_rect_d3", "ga1_rect_d5_tw2"]


def _run(code, env_extra):
    env = dict(os.environ, **env_extra)
    pre = "import sys; sys.path[:0] = [%r, %r]\n" % (ROOT, os.path.join(ROOT, "tests"))
    return subprocess.run([sys.executable, "-c", pre + code], capture_output=True, text=True, env=env, timeout=600)


def test_library_variants_name_their_arithmetic(lib):
    """Both builds export the whole C ABI (capi.load checks every symbol) and say which arithmetic their
    render kernels use; the loader refuses a library that is not what its name says."""
    assert lib.eu_render_arithmetic() == 0
    r = _run("import os\nfrom envutil_b200 import capi\nlib = capi.load()\n"
             "print(capi.ARITHMETIC, os.path.basename(capi.LIB_PATH), lib.eu_render_arithmetic())\n",
             {"EU_ARITHMETIC": "contracted"})
    assert r.returncode == 0, r.stderr[-1500:]
    assert r.stdout.split() == ["contracted", "libenvutil_b200_fma.so", "1"]


def test_variants_share_everything_but_the_render_kernels():
    """Staging, API and host set-up are the same objects in both libraries (so a staged source has the same
    bits in both); only the render translation units are compiled twice."""
    import __graft_entry__ as g
    assert set(g.FMA_SOURCES) == {s for s in g.CUDA_SOURCES if s.startswith("render_")}
    assert "stage.cu" not in g.FMA_SOURCES and "api.cu" not in g.FMA_SOURCES and "render.cu" not in g.FMA_SOURCES


@pytest.mark.parametrize("name", SUBSET)
def test_contracted_arithmetic_stays_within_tolerance(name):
    job = jobs.JOBS[name]
    exact, idx = harness.oracle_render(job, want_index=True)
    fused, fidx = harness.oracle_render(job, want_index=True, contracted=True)
    assert np.array_equal(idx, fidx)  # rays, coordinates, gates, faces and facets are untouched
    c = harness.compare(fused, exact)
    assert c["max_rel"] <= 1e-5 and c["rms_rel"] <= 5e-7, c
    # ... and the switch does not leak: the next render is the pinned one again
    assert np.array_equal(harness.oracle_render(job), exact)


def test_contracted_arithmetic_differs_where_it_should():
    """A cubic job: the variant is not a no-op (some values move by an ulp or two), a degree-0 job
    (nearest neighbour, no arithmetic in the window) is untouched."""
    job = jobs.JOBS["ll_rect_d3_rot"]
    a, b = harness.oracle_render(job), harness.oracle_render(job, contracted=True)
    assert 0 < int((a != b).sum())
    near = [n for n in jobs.JOBS if jobs.JOBS[n].degree == 0 and not jobs.JOBS[n].twine]
    for n in near[:1]:
        assert np.array_equal(harness.oracle_render(jobs.JOBS[n]), harness.oracle_render(jobs.JOBS[n], contracted=True))


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("EU_GPU_UNTRIED") != "1",
                    reason="no GPU run has seen this build yet: opt in with EU_GPU_UNTRIED=1 (tools/gpu_first_call.sh does)")
def test_contracted_kernels_equal_contracted_oracle():
    """All small jobs through libenvutil_b200_fma.so in ONE separate process (the library is a process-wide
    singleton): bit-identical to the oracle's restatement of the variant, within tolerance of the pinned one."""
    code = ("import numpy as np, harness, jobs\n"
            "from envutil_b200 import capi\n"
            "from envutil_b200.engine import Engine\n"
            "assert capi.load().eu_render_arithmetic() == 1\n"
            "eng = Engine(0)\n"
            "bad = 0; worst = 0.0\n"
            "for name in sorted(jobs.JOBS):\n"
            "    job = jobs.JOBS[name]\n"
            "    out = eng.render(job)\n"
            "    c = harness.compare(out, harness.oracle_render(job, contracted=True))\n"
            "    e = harness.compare(out, harness.oracle_render(job))\n"
            "    worst = max(worst, e['max_rel'])\n"
            "    if c['n_diff'] or e['max_rel'] > (5e-5 if 'hdr' in name else 1e-5):\n"
            "        bad += 1; print(name, c, e, flush=True)\n"
            "print('worst max_rel vs the pinned arithmetic', worst)\n"
            "eng.close()\n"
            "sys.exit(3 if bad else 0)\n")
    r = _run(code, {"EU_ARITHMETIC": "contracted"})
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]


def test_contracted_oracle_is_pinned_to_itself():
    """The variant's restatement has no reference build to be pinned to (a compiler chooses its own
    contractions), so its outputs are pinned to themselves: tests/golden/contracted.json holds the hashes of
    the twelve jobs above as first written, against accidental changes of the restatement."""
    import hashlib
    import json
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "contracted.json")))
    assert sorted(want) == sorted(SUBSET)
    for name in SUBSET:
        out = harness.oracle_render(jobs.JOBS[name], contracted=True)
        assert list(out.shape) == want[name]["shape"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == want[name]["sha256"], name
