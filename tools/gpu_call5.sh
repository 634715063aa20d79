#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-call5}
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -q -m gpu -s
step bench 900 python bench.py
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
step bench_contracted 900 python bench.py --arithmetic contracted --no-cpu-baseline
tail -n 1 "$OUT/bench_contracted.log" > "$OUT/bench_contracted.json"
step calibrate 300 python tools/calibrate_c5_cost.py
step reference_c2 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
NCU="ncu --set full --clock-control none --import-source on"
step ncu_c2 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C2" python tools/bench_configs.py --configs C2 --steps 3 --padded 32
step ncu_c5a 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C5A" python tools/bench_configs.py --configs C5 --steps 3 --padded 32
step ncu_c5b 500 $NCU -k regex:k_render -s 39 -c 1 -o "$OUT/prof_C5B" python tools/bench_configs.py --configs C5 --steps 3 --padded 32
export EU_PROFILE_DIR="$OUT/profiles"
step summarise 300 python tools/summarise_kernels.py r02 C2="$OUT/prof_C2.ncu-rep" C5A="$OUT/prof_C5A.ncu-rep" C5B="$OUT/prof_C5B.ncu-rep"
rm -f "$OUT"/prof_C5A.ncu-rep "$OUT"/prof_C5B.ncu-rep
cat "$OUT/summary.txt"
