#!/bin/bash
# One GPU: the whole GPU test suite, the bench line, the upload/download probe, an ncu capture of the C3a kernel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-call9}
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -q -m gpu -s
step bench 900 python bench.py
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"

step bench_c5_n1 600 python bench.py --workload c5 --steps 20 --warmup 5
grep '^{"metric"' "$OUT/bench_c5_n1.log" | tail -n 1 > "$OUT/bench_c5_n1.json"
NCU="ncu --set full --clock-control none --import-source on"
step ncu_c3b 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C3b" python tools/bench_configs.py --configs C3b --steps 3 --padded 32
step ncu_c5b 500 $NCU -k regex:k_render -s 39 -c 1 -o "$OUT/prof_C5B" python tools/bench_configs.py --configs C5 --steps 3 --padded 32
export EU_PROFILE_DIR="$OUT/profiles"
step summarise 300 python tools/summarise_kernels.py r02 C3b="$OUT/prof_C3b.ncu-rep" C5B="$OUT/prof_C5B.ncu-rep"
cat "$OUT/summary.txt"
tail -n 3 "$OUT/pytest_gpu.log"
