"""The drop-in boundary: libenvutil_b200.so loads, exports every symbol include/envutil_b200.h
declares, keeps the struct layout the header states, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from envutil_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "envutil_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eu_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in capi.SYMBOLS, f"{n} is declared in the header but not bound in capi.py"
    assert sorted(capi.SYMBOLS) == names


def test_struct_layout_matches_header(lib, tmp_path):
    """sizeof/offsetof as the C compiler sees the header == the ctypes mirror."""
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "envutil_b200.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(eu_facet_t), sizeof(eu_target_t),'
                    'sizeof(eu_opts_t), sizeof(eu_tap_t), sizeof(eu_timing_t), offsetof(eu_facet_t, brighten),'
                    'offsetof(eu_facet_t, shift_h), offsetof(eu_facet_t, window_width), offsetof(eu_target_t, crop_x0),'
                    'offsetof(eu_target_t, single));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.Facet), C.sizeof(capi.Target), C.sizeof(capi.Opts), C.sizeof(capi.Tap),
            C.sizeof(capi.Timing), capi.Facet.brighten.offset, capi.Facet.shift_h.offset,
            capi.Facet.window_width.offset, capi.Target.crop_x0.offset, capi.Target.single.offset]
    assert got == want


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c89.c"
    prog.write_text('#include "envutil_b200.h"\nint main(void){return 0;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           str(prog), "-o", str(tmp_path / "c89.o")])


def test_no_cpu_fallback(lib):
    """Without a device the library refuses to initialise, and every device entry point reports
    EU_ERR_STATE before eu_init: nothing renders on the CPU."""
    if lib.eu_device_count() > 0:
        pytest.skip("a GPU is present")
    assert lib.eu_init(0) == -4  # EU_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.eu_last_error()
    t, o = capi.Target(), capi.Opts()
    assert lib.eu_render(C.byref(t), C.byref(o), 0, None, None, None, 0, None, None) == -5  # EU_ERR_STATE
    assert lib.eu_cycle() == -5
    assert lib.eu_source_find(b"x") is None
    f, h, ptr, pitch = capi.Facet(), capi.SourceH(), C.c_void_p(), C.c_int()
    assert lib.eu_source_reserve(None, C.byref(f), C.byref(o), C.byref(h), C.byref(ptr), C.byref(pitch), C.byref(pitch)) == -5
    assert lib.eu_render_rows_pitched(C.byref(t), C.byref(o), 0, None, None, None, 0, 0, 1, None, 0, None, None) == -5
    assert lib.eu_frame_alloc(16, C.byref(ptr)) == -5
    assert lib.eu_frame_open(b"\0" * 64, C.byref(ptr)) == -5


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: no product file may name it."""
    pkg = os.path.join(ROOT, "envutil_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "liboracle" not in txt and "eu_oracle" not in txt and "orc_" not in txt, os.path.join(d, f)
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
