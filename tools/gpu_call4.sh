#!/bin/bash
# One GPU: GPU suite, then A/B of the footprint-staged kernel (bit 4 = previous kernel) and the texel layouts.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-call4}
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -q -m gpu -s
step fuzz_gpu 400 python tools/fuzz_gpu_vs_oracle.py --n 300 --seed 31
step c2_ab 300 python tools/bench_configs.py --configs C2 --padded 0,16,8,24,1,17,0,16 --steps 20
step variants 600 python tools/bench_configs.py --configs C1,C3a,C3b,C4 --padded 32,40,0,1 --steps 10
step bench 900 python bench.py
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
NCU="ncu --set full --clock-control none --import-source on"
step ncu_c2 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C2" python tools/bench_configs.py --configs C2 --steps 3
step ncu_c4 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C4" python tools/bench_configs.py --configs C4 --steps 3 --padded 32
step ncu_c3a 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C3a" python tools/bench_configs.py --configs C3a --steps 3 --padded 32
step ncu_c3b 600 $NCU -k regex:k_render -s 9 -c 1 -o "$OUT/prof_C3b" python tools/bench_configs.py --configs C3a,C3b --steps 3 --padded 32
step ncu_c1 300 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C1" python tools/bench_configs.py --configs C1 --steps 3 --padded 32
export EU_PROFILE_DIR="$OUT/profiles"
step summarise 300 python tools/summarise_kernels.py r02 C1="$OUT/prof_C1.ncu-rep" C2="$OUT/prof_C2.ncu-rep" C3a="$OUT/prof_C3a.ncu-rep" C3b="$OUT/prof_C3b.ncu-rep" C4="$OUT/prof_C4.ncu-rep"
rm -f "$OUT"/prof_C1.ncu-rep "$OUT"/prof_C3a.ncu-rep "$OUT"/prof_C3b.ncu-rep
cat "$OUT/summary.txt"
