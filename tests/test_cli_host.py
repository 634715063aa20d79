"""The C++ host (envutil_b200_cli): envutil's command line and PTO subset restated above the C
ABI. CPU part: --dry_run prints what the kernels would be given; it must equal the marshalling
the parity tests use (which is pinned to the reference binary through the golden outputs)."""
import os
import re
import subprocess

import numpy as np
import pytest

import harness
import jobs
from envutil_b200 import euf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "envutil_b200", "envutil_b200_cli")


@pytest.fixture(scope="module")
def cli(lib):
    if not os.path.exists(CLI):
        import __graft_entry__ as g
        g.build_cli()
    return CLI


def _write_facets(job, d):
    paths = []
    for i, f in enumerate(job.facets):
        p = os.path.join(d, "facet%d.euf" % i)
        euf.write_euf(p, f.image)
        paths.append(p)
    return paths


def _floats(line, key):
    m = re.search(key + r" ((?:[-+0-9.eEinfa]+ ?)+)", line)
    return [float(v) for v in m.group(1).split()]


@pytest.mark.parametrize("name", ["ll_rect_d1_oddangles", "ll_fish_d1_tw4", "ll_rect_d1_tw3_sigma", "hdr3_sph_d3_tw2",
                                  "lens3_voronoi_sph_d1", "eev_voronoi_sph_d1", "ll_cube_d3_rot", "voronoi4_solo2",
                                  "cm_sph_d1_support4_tile16", "cropout_ll_rect_d3_tw2", "cropout_voronoi4_fish_d1"])
def test_dry_run_equals_marshalling(cli, tmp_path, name):
    job = jobs.JOBS[name]
    paths = _write_facets(job, str(tmp_path))
    r = subprocess.run([cli] + job.cli_args(paths, str(tmp_path / "out.euf")) + ["--dry_run"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    t, fa, o, taps, ntaps = job.structs()
    tl = [l for l in lines if l.startswith("target ")][0]
    assert "%dx%d" % (t.width, t.height) in tl
    if t.crop_width > 0:  # p-line S clause
        assert "crop %dx%d+%d+%d" % (t.crop_width, t.crop_height, t.crop_x0, t.crop_y0) in lines
    assert _floats(tl, "hfov")[0] == t.hfov and _floats(tl, "yaw")[0] == t.yaw
    assert _floats(tl, "pitch")[0] == t.pitch and _floats(tl, "roll")[0] == t.roll
    el = [l for l in lines if l.startswith("extent ")][0]
    assert _floats(el, "extent") == [t.x0, t.x1, t.y0, t.y1] and _floats(el, "step")[0] == t.step
    fl = [l for l in lines if l.startswith("facet ")]
    assert len(fl) == len(job.facets)
    for i, l in enumerate(fl):
        assert _floats(l, "hfov")[0] == fa[i].hfov
        assert _floats(l, "ypr") == [fa[i].yaw, fa[i].pitch, fa[i].roll]
        assert _floats(l, " step")[0] == fa[i].step
        assert np.float32(_floats(l, "brighten")[0]) == np.float32(fa[i].brighten)
        assert _floats(l, "shift") == [fa[i].shift_h, fa[i].shift_v]
        assert _floats(l, "shear") == [fa[i].shear_g, fa[i].shear_t]
    tp = [l for l in lines if l.startswith("tap ")]
    assert len(tp) == ntaps
    for k, l in enumerate(tp):
        x, y, w = (np.float32(v) for v in l.split()[1:])
        assert (x, y, w) == (np.float32(taps[k].x), np.float32(taps[k].y), np.float32(taps[k].w))


def test_input_alias_and_spline_degree(cli, tmp_path):
    """--input X == one facet with the projection inferred from the aspect
    (2:1 -> spherical 360, 1:6 -> cubemap 90); --spline_degree == --degree."""
    job = jobs.JOBS["cm_sph_d3"]
    p = _write_facets(job, str(tmp_path))[0]
    base = ["--projection", "spherical", "--hfov", "360", "--width", "192", "--height", "96", "--twine", "0",
            "--output", str(tmp_path / "o.euf"), "--dry_run"]
    a = subprocess.run([cli, "--input", p, "--spline_degree", "3"] + base, capture_output=True, text=True)
    b = subprocess.run([cli, "--facet", p, "cubemap", "90", "0", "0", "0", "--degree", "3"] + base,
                       capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout == b.stdout and "degree 3" in a.stdout


def test_photo_is_a_rectilinear_65_degree_facet(cli, tmp_path):
    """--photo IMAGE (envutil_main.cc:916-927): projection and hfov come from image metadata; without any
    (.euf has none) the reference assumes rectilinear, 65 degrees. Photos are numbered after the facets."""
    job = jobs.JOBS["rect_src_sph_d1"]
    p = _write_facets(job, str(tmp_path))[0]
    base = ["--projection", "spherical", "--hfov", "360", "--width", "128", "--height", "64", "--twine", "0",
            "--output", str(tmp_path / "o.euf"), "--dry_run"]
    a = subprocess.run([cli, "--photo", p] + base, capture_output=True, text=True)
    b = subprocess.run([cli, "--facet", p, "rectilinear", "65", "0", "0", "0"] + base, capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout == b.stdout and "rectilinear" in a.stdout
    c = subprocess.run([cli, "--photo", p, "--facet", p, "rectilinear", "80", "30", "0", "0"] + base, capture_output=True,
                       text=True)
    fl = [l for l in c.stdout.splitlines() if l.startswith("facet ")]
    assert c.returncode == 0 and len(fl) == 2
    assert _floats(fl[0], "hfov")[0] == 80 * (np.pi / 180.0) and _floats(fl[1], "hfov")[0] == 65 * (np.pi / 180.0)


def test_pano_clause_is_a_solo_panorama_facet(cli, tmp_path):
    """i-line 'Pano' clause (envutil_main.cc:673-712, 'unstitching'): the facet takes the p-line's projection,
    hfov (and crop) and becomes the solo facet that --split re-creates the others from. The job equals the
    one spelled out with an ordinary i-line, --solo and an explicit target (the reference binary gives
    bit-identical files for the two forms; checked when this test was written)."""
    pano = np.zeros((64, 128, 3), np.float32)
    euf.write_euf(str(tmp_path / "pano.euf"), pano)
    for n in ("a", "b"):
        euf.write_euf(str(tmp_path / (n + ".euf")), np.zeros((64, 96, 3), np.float32))
    rest = ('i w96 h64 f0 v80 y30 p10 r5 a0.01 b-0.02 n"%s"\ni w96 h64 f0 v70 y-40 p-5 r0 TrX0.02 n"%s"\n'
            % (tmp_path / "a.euf", tmp_path / "b.euf"))
    (tmp_path / "A.pto").write_text('p f2 w128 h64 v360\ni w128 h64 f4 v360 y0 p0 r0 Pano"%s"\n' % (tmp_path / "pano.euf") + rest)
    (tmp_path / "B.pto").write_text('i w128 h64 f4 v360 y0 p0 r0 n"%s"\n' % (tmp_path / "pano.euf") + rest)
    common = ["--twine", "0", "--split", str(tmp_path / "re%02d.euf"), "--dry_run"]
    a = subprocess.run([cli, "--pto", str(tmp_path / "A.pto")] + common, capture_output=True, text=True)
    b = subprocess.run([cli, "--pto", str(tmp_path / "B.pto"), "--solo", "0", "--projection", "spherical", "--hfov", "360",
                        "--width", "128", "--height", "64"] + common, capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout == b.stdout and "solo 0" in a.stdout
    # cropped panorama: the file holds the p-line's S window
    euf.write_euf(str(tmp_path / "crop.euf"), np.zeros((40, 100, 3), np.float32))
    (tmp_path / "C.pto").write_text('p f2 w128 h64 v360 S10,110,12,52\ni w128 h64 f4 v360 y0 p0 r0 Pano"%s"\n'
                                    % (tmp_path / "crop.euf") + rest)
    c = subprocess.run([cli, "--pto", str(tmp_path / "C.pto")] + common, capture_output=True, text=True)
    assert c.returncode == 0, c.stderr
    assert re.search(r"facet 0 \S+ spherical 128x64x3", c.stdout)
    bad = subprocess.run([cli, "--pto_line", 'i w128 h64 f4 v360 Pano"%s"' % (tmp_path / "pano.euf"), "--output", "x.euf",
                          "--dry_run"], capture_output=True, text=True)
    assert bad.returncode != 0 and "needs a p-line" in bad.stderr


def test_pipe_mode_and_errors(cli, tmp_path):
    job = jobs.JOBS["ll_rect_d1"]
    p = _write_facets(job, str(tmp_path))[0]
    common = [cli, "--facet", p, "spherical", "360", "0", "0", "0", "--dry_run", "-"]
    feed = ("--projection rectilinear --hfov 90 --width 96 --height 54 --output '%s'\n"
            "--projection fisheye --hfov 180 --width 64 --height 64 --yaw 10 --output \"%s\"\n"
            % (tmp_path / "a b.euf", tmp_path / "b.euf"))
    r = subprocess.run(common, input=feed, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("target ") == 2 and "target fisheye 64x64" in r.stdout and "pipe has reached EOF" in r.stdout
    bad = subprocess.run([cli, "--facet", p, "spherical", "360", "0", "0", "0", "--frobnicate", "1", "--output", "x.euf"],
                         capture_output=True, text=True)
    assert bad.returncode != 0 and "unknown argument" in bad.stderr
    bad = subprocess.run([cli, "--facet", p, "spherical", "360", "0", "0", "0", "--projection", "cubemap", "--hfov", "60",
                          "--output", "x.euf", "--dry_run"], capture_output=True, text=True)
    assert bad.returncode != 0 and "hfov >= 90" in bad.stderr
    bad = subprocess.run([cli, "--facet", str(tmp_path / "missing.euf"), "spherical", "360", "0", "0", "0", "--output",
                          "x.euf"], capture_output=True, text=True)
    assert bad.returncode != 0 and "failed to open facet image" in bad.stderr


def test_pto_back_references(cli, tmp_path):
    """`=N` takes the field from i-line N (reference pto.h:137-147)."""
    job = jobs.JOBS["voronoi3_rect_d1_tw2"]
    paths = _write_facets(job, str(tmp_path))
    pto = tmp_path / "p.pto"
    pto.write_text("# comment\np f2 w200 h100 v360\n"
                   'i w96 h64 f0 v80 y-100 p0 r0 Eev12 n"%s"\n'
                   'i w96 h64 f0 v=0 y-30 p12 r3 Eev=0 n"%s"\n'
                   'i w96 h64 f0 v=0 y40 p-9 r-5 Eev14 n"%s"\n' % tuple(paths))
    r = subprocess.run([cli, "--pto", str(pto), "--output", str(tmp_path / "o.euf"), "--dry_run"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    assert "target spherical 200x100" in r.stdout
    fl = [l for l in r.stdout.splitlines() if l.startswith("facet ")]
    hf = [_floats(l, "hfov")[0] for l in fl]
    assert hf[0] == hf[1] == hf[2] == 80 * (np.pi / 180.0)
    br = [_floats(l, "brighten")[0] for l in fl]
    mean = np.float32(np.float32(12 + 12 + 14) / np.float32(3))
    want = [np.float32(2.0 ** float(np.float32(np.float32(e) - mean))) for e in (12, 12, 14)]
    assert [np.float32(b) for b in br] == want


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ll_rect_d1_oddangles", "cm_sph_d3_rot", "ll_fish_d1_tw4", "hdr3_sph_d3_tw2",
                                  "lens3_voronoi_sph_d1", "eev_voronoi_sph_d1", "ll_ba6_d1_tw2", "voronoi4_solo2",
                                  "grey_cm_sph_d1_tw2", "ll_cyl_d1_tw2", "tr3_voronoi_fish_d3_tw2",
                                  "mixed_grey_rgba_voronoi_sph_d1", "rgba4_voronoi_sph_d3", "hdr3_rgba_rect_d1",
                                  "auto_tw_voronoi_d3", "auto_tw_up_ll_rect_d1", "auto_tw_density_ll_ba6",
                                  "mask_crop4_voronoi_sph_d1", "crop_fish_sph_d1", "mask_grey_sph_d1_tw2",
                                  "win_voronoi_sph_d1", "win_rect_rect_d3_tw2", "single1_hdr3_d1",
                                  "single2_voronoi4_d3_tw2", "single0_cm_ll", "single0_lens3_d1",
                                  "single1_lens3_d3_tw2", "single1_tr3_d1_tw2", "single1_tr_lens_d1", "cropout_ll_sph_d1",
                                  "cropout_ll_rect_d3_tw2", "cropout_ll_cyl_d1_tw2", "cropout_voronoi4_fish_d1",
                                  "maskfor2_voronoi4_grey_d3_tw2", "maskfor3_rgba4_ga_rect_d1", "maskfor1_mixed_rgb_rgba_d1",
                                  "nch1_voronoi4_sph_d1", "nch4_grey_ll_rect_d3", "nch3_rgba1_rect_d1_tw2"])
def test_cli_output_equals_reference_output(cli, tmp_path, name):
    """The drop-in claim end to end: the SAME command line given to the reference binary and to
    envutil_b200_cli produces the same file, bit for bit (golden sha256 of the reference run)."""
    import hashlib
    import json
    job = jobs.JOBS[name]
    paths = _write_facets(job, str(tmp_path))
    outp = str(tmp_path / "out.euf")
    r = subprocess.run([cli] + job.cli_args(paths, outp), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    img = euf.read_euf(outp)
    man = json.load(open(os.path.join(harness.GOLDEN, "manifest.json")))[name]
    assert list(img.shape) == man["shape"]
    assert hashlib.sha256(np.ascontiguousarray(img, dtype="<f4").tobytes()).hexdigest() == man["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cm_sph_d3", "ga2_voronoi_sph_d1", "single1_hdr3_d1"])
def test_cli_tethered_frame_equals_reference_frame(cli, tmp_path, name):
    """The same command line, rendered tethered (--screen_out: payload() stores uint32 sRGBA as it would into visor's
    frame buffer), equals the frame the reference's own tethered pipeline produced (tests/golden/screen.json)."""
    import hashlib
    import json
    job = jobs.JOBS[name]
    paths = _write_facets(job, str(tmp_path))
    scr = str(tmp_path / "frame.u32")
    r = subprocess.run([cli] + job.cli_args(paths, str(tmp_path / "none.euf")) + ["--screen_out", scr], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    man = json.load(open(os.path.join(harness.GOLDEN, "screen.json")))[name]
    frame = np.fromfile(scr, dtype="<u4")
    assert frame.size == man["shape"][0] * man["shape"][1]
    assert hashlib.sha256(frame.tobytes()).hexdigest() == man["sha256"]
    assert not os.path.exists(str(tmp_path / "none.euf"))  # tethered jobs write no image file


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(jobs.CLI_EXTRAS))
def test_cli_twf_filter_equals_reference(cli, tmp_path, name):
    """--twf_file / --twine_normalize / --twine_width scaling: same command line, same bits."""
    import hashlib
    import json
    base, twf, extra = jobs.CLI_EXTRAS[name]
    job = jobs.JOBS[base]
    paths = _write_facets(job, str(tmp_path))
    tp = tmp_path / "filter.twf"
    tp.write_text(twf)
    outp = str(tmp_path / "out.euf")
    r = subprocess.run([cli] + job.cli_args(paths, outp) + ["--twf_file", str(tp)] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    img = euf.read_euf(outp)
    man = json.load(open(os.path.join(harness.GOLDEN, "manifest.json")))[name]
    assert hashlib.sha256(np.ascontiguousarray(img, dtype="<f4").tobytes()).hexdigest() == man["sha256"]


FACES = ["left", "right", "top", "bottom", "front", "back"]


def test_cubeface_series_input_is_one_cubemap(cli, tmp_path):
    """A facet name with one '%' is a cubeface series (envutil_basic.h:267-356): six square images named
    by direction stand for the 1:6 stripe. Dry run: the job is the same as for the stripe in one file."""
    job = jobs.JOBS["cm_sph_d3"]
    img = job.facets[0].image
    F = img.shape[1]
    for i, n in enumerate(FACES):
        euf.write_euf(str(tmp_path / ("face_%s.euf" % n)), np.ascontiguousarray(img[i * F:(i + 1) * F]))
    one = _write_facets(job, str(tmp_path))
    a = subprocess.run([cli] + job.cli_args([str(tmp_path / "face_%s.euf")], "o.euf") + ["--dry_run"], capture_output=True,
                       text=True)
    b = subprocess.run([cli] + job.cli_args(one, "o.euf") + ["--dry_run"], capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    strip = lambda out: [re.sub(r"^(facet \d+) \S+", r"\1", l) for l in out.splitlines()]  # the file name differs
    assert strip(a.stdout) == strip(b.stdout)


@pytest.mark.gpu
def test_cli_cubeface_series_in_and_out(cli, tmp_path):
    """Six-file cubemaps on both sides: reading a series gives the reference's output for the stripe
    (golden), writing a cubemap to a '%' name gives six files that are the faces of the stripe."""
    import hashlib
    import json
    man = json.load(open(os.path.join(harness.GOLDEN, "manifest.json")))
    job = jobs.JOBS["cm_sph_d3"]
    img = job.facets[0].image
    F = img.shape[1]
    for i, n in enumerate(FACES):
        euf.write_euf(str(tmp_path / ("face_%s.euf" % n)), np.ascontiguousarray(img[i * F:(i + 1) * F]))
    outp = str(tmp_path / "o.euf")
    r = subprocess.run([cli] + job.cli_args([str(tmp_path / "face_%s.euf")], outp), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = euf.read_euf(outp)
    assert hashlib.sha256(np.ascontiguousarray(got, dtype="<f4").tobytes()).hexdigest() == man["cm_sph_d3"]["sha256"]
    job = jobs.JOBS["ll_cube_d1"]
    p = _write_facets(job, str(tmp_path))
    r = subprocess.run([cli] + job.cli_args(p, str(tmp_path / "out_%s.euf")), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    faces = [euf.read_euf(str(tmp_path / ("out_%s.euf" % n))) for n in FACES]
    whole = np.concatenate(faces, axis=0)
    assert list(whole.shape) == man["ll_cube_d1"]["shape"]
    assert hashlib.sha256(np.ascontiguousarray(whole, dtype="<f4").tobytes()).hexdigest() == man["ll_cube_d1"]["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(jobs.SPLITS))
def test_cli_split_equals_reference(cli, tmp_path, name):
    """--split FORMAT: one 'single' job per facet (the solo facet excepted), each through the inverse
    lens / translation path where the facet has one; every file equals the reference's."""
    import hashlib
    import json
    man = json.load(open(os.path.join(harness.GOLDEN, "manifest.json")))[name]
    base, extra = jobs.SPLITS[name]
    job = jobs.JOBS[base]
    paths = _write_facets(job, str(tmp_path))
    args = job.cli_args(paths, "unused.euf")
    k = args.index("--output")
    del args[k:k + 2]
    r = subprocess.run([cli] + args + extra + ["--split", str(tmp_path / "re%02d.euf")], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    made = sorted(p.name for p in tmp_path.glob("re*.euf"))
    assert made == sorted("re%02d.euf" % int(i) for i in man["outputs"])
    for i, want in man["outputs"].items():
        img = euf.read_euf(str(tmp_path / ("re%02d.euf" % int(i))))
        assert list(img.shape) == want["shape"]
        assert hashlib.sha256(np.ascontiguousarray(img, dtype="<f4").tobytes()).hexdigest() == want["sha256"], i


@pytest.mark.gpu
def test_cli_pipe_mode_keeps_sources_staged(cli, tmp_path):
    job = jobs.JOBS["cm_sph_d3"]
    p = _write_facets(job, str(tmp_path))[0]
    feed = "".join("--yaw %d --output %s\n" % (y, tmp_path / ("o%d.euf" % y)) for y in (0, 40, 80))
    r = subprocess.run([cli, "-v", "--facet", p, "cubemap", "90", "0", "0", "0", "--projection", "spherical", "--hfov",
                        "360", "--width", "192", "--height", "96", "--degree", "3", "--twine", "0", "-"], input=feed,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    st = [float(m) for m in re.findall(r"staging ([0-9.]+) ms", r.stdout)]
    assert len(st) == 3 and st[0] > 0.0 and st[1] == 0.0 and st[2] == 0.0  # staged once, found twice
    a = euf.read_euf(str(tmp_path / "o0.euf"))
    assert np.array_equal(a, harness.oracle_render(job))
