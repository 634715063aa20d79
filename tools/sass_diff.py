#!/usr/bin/env python3
"""Which kernels changed? Compiles the CUDA sources of a git revision into a temporary directory with the
flags of __graft_entry__.build() and compares them, kernel by kernel, with the objects of the current build
(envutil_b200/csrc/build/*.o): the SASS of every kernel the revision had, addresses and encodings stripped of
nothing - a kernel counts as unchanged only if its instruction listing is byte-identical.

Used at the end of round 1, when no GPU time was left, to add opt-in code (the FMA-window build, k_render_warp)
while proving that the kernels the GPU tests had last passed with were still exactly the ones in the library:

  python tools/sass_diff.py fcee592        # 232 kernels at that revision, 0 changed, 19 new
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def kernels(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            out[cur].append(line.strip())
    return {k: hashlib.sha1("\n".join(v).encode()).hexdigest() for k, v in out.items()}


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
    g.build_library()
    with tempfile.TemporaryDirectory(prefix="sassdiff_") as d:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "envutil_b200/csrc", "include"], capture_output=True, check=True)
        subprocess.run(["tar", "-x", "-C", d], input=tar.stdout, check=True)
        csrc = os.path.join(d, "envutil_b200", "csrc")
        srcs = [s for s in g.CUDA_SOURCES if s.endswith(".cu") and os.path.exists(os.path.join(csrc, s))]
        flags = [f for f in g.NVCC_FLAGS if not f.startswith("-I")] + ["-I" + os.path.join(d, "include"), "-I" + csrc]

        def cc(src):
            o = os.path.join(d, src[:-3] + ".o")
            g._run([g.NVCC] + flags + ["-c", os.path.join(csrc, src), "-o", o])
            return src, o
        with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
            objs = dict(ex.map(cc, srcs))
        total = changed = new = gone = 0
        for src in srcs:
            old = kernels(objs[src])
            cur_obj = os.path.join(g.OBJ, src[:-3] + ".o")
            cur = kernels(cur_obj) if os.path.exists(cur_obj) else {}
            total += len(old)
            for k in old:
                if k not in cur:
                    gone += 1
                    print("gone   ", src, demangle(k)[:110])
                elif cur[k] != old[k]:
                    changed += 1
                    print("CHANGED", src, demangle(k)[:110])
            for k in cur:
                if k not in old:
                    new += 1
                    print("new    ", src, demangle(k)[:110])
        print("%d kernels at %s: %d changed, %d gone; %d new in the current build" % (total, rev, changed, gone, new))
        return 1 if changed or gone else 0


if __name__ == "__main__":
    sys.exit(main())
