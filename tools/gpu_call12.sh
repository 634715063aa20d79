#!/bin/bash
# One GPU: the whole GPU test suite and the bench line (after --mask_for).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-call12}
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -q -m gpu -s
step fuzz 600 python tools/fuzz_gpu_vs_oracle.py --n 300 --seed 21
step bench 900 python bench.py
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
cat "$OUT/summary.txt"
tail -n 3 "$OUT/pytest_gpu.log"; tail -n 3 "$OUT/fuzz.log"
