"""The contracted arithmetic (EU_OPT_CONTRACTED in eu_opts_t.reserved[1], include/envutil_b200.h): fused
multiply-adds in the b-spline window evaluation and the twining accumulation, nothing else. The default arithmetic
stays bit-identical to the reference's parity build; this one is what a reference compiled with g++'s default
-ffp-contract=fast on FMA hardware resembles, and it has to stay within BASELINE.json's tolerance (1e-5 relative)
of the pinned reference. Both sets of kernels live in the one library.

CPU: the oracle's restatement of the variant (orc_set_arithmetic) keeps every index plane and differs from the
pinned arithmetic by at most 2.6e-6 relative (hdr_merge's near-zero denominators: 1.6e-5), RMS <= 2.7e-7, over all
98 small jobs. GPU: the contracted kernels equal the contracted oracle bit for bit on every small job.
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


def test_library_carries_both_arithmetics(lib):
    """One library, two sets of render kernels; Job.contracted / EU_ARITHMETIC=contracted set the option bit."""
    assert lib.eu_render_arithmetic() == 2
    import copy
    job = copy.copy(jobs.JOBS["ll_rect_d3_rot"])
    assert job.structs(lib)[2].reserved[1] == 0
    job.contracted = True
    assert job.structs(lib)[2].reserved[1] == 16
    r = _run("import jobs\nfrom envutil_b200 import capi\n"
             "print(capi.ARITHMETIC, jobs.JOBS['ll_rect_d3_rot'].structs()[2].reserved[1])\n",
             {"EU_ARITHMETIC": "contracted"})
    assert r.returncode == 0, r.stderr[-1500:]
    assert r.stdout.split() == ["contracted", "16"]


def test_variants_share_everything_but_the_render_kernels():
    """Staging, API and host set-up are compiled once (so a staged source has the same bits for both arithmetics);
    only the render translation units are compiled twice."""
    import __graft_entry__ as g
    assert set(g.FMA_SOURCES) == {s for s in g.CUDA_SOURCES if s.startswith("render_") and s != "render_tie.cu"}
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
def test_contracted_kernels_equal_contracted_oracle(engine):
    """Every small job with EU_OPT_CONTRACTED: bit-identical to the oracle's restatement of the variant, within
    tolerance of the pinned arithmetic, and the option does not leak into the next (exact) job."""
    import copy
    worst = 0.0
    for name in sorted(jobs.JOBS):
        job = copy.copy(jobs.JOBS[name])
        job.contracted = True
        out = engine.render(job)
        c = harness.compare(out, harness.oracle_render(job, contracted=True))
        e = harness.compare(out, harness.oracle_render(job))
        worst = max(worst, e["max_rel"])
        assert c["n_diff"] == 0, (name, c)
        assert e["max_rel"] <= (5e-5 if "hdr" in name else 1e-5), (name, e)
    print("worst max_rel vs the pinned arithmetic", worst)
    job = jobs.JOBS["ll_rect_d3_rot"]
    assert harness.compare(engine.render(job), harness.oracle_render(job))["n_diff"] == 0


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
