#!/usr/bin/env python3
"""Generate tests/golden/screen.json (+ screen_<job>.npz for two small jobs): the uint32 sRGBA frames the UNMODIFIED
reference writes into its viewer's frame buffer when it runs tethered (handle_job -> core(tethered) -> work():
act + to_screen_t, envutil_main.cc:1755-1868, envutil_payload.cc:298-413,524-531), for tests/jobs.py SCREEN_JOBS.
The oracle build's visor stub (oracle/shim/visor_stub/visor.h) hands the reference one job and saves the buffer.
Only this container can run it (it needs oracle/_ref, built from /root/reference); tests read the committed files."""
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

FULL = ["ll_rect_d1", "rgba1_rect_d1"]


def main():
    man = {}
    for name in jobs.SCREEN_JOBS:
        job = jobs.JOBS[name]
        pm = harness.reference_screen(job, "pm")
        lm = harness.reference_screen(job, "libm")
        a, b = pm.view(np.uint8).astype(int), lm.view(np.uint8).astype(int)
        man[name] = {"shape": list(pm.shape), "sha256": hashlib.sha256(np.ascontiguousarray(pm, dtype="<u4").tobytes()).hexdigest(),
                     "libm_vs_pinned": {"bytes_differing": int((a != b).sum()), "max_step": int(np.abs(a - b).max())}}
        if name in FULL:
            np.savez_compressed(os.path.join(harness.GOLDEN, "screen_" + name + ".npz"), out=pm)
        print(name, man[name])
    json.dump(man, open(os.path.join(harness.GOLDEN, "screen.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
