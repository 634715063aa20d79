#!/bin/bash
# One GPU, the state of HEAD: the whole GPU suite, the kernels-vs-oracle sweep, the bench line, one ncu --set full capture
# of every config's render kernel (summarised into $OUT/profiles/r02_kernels.json + r02_<config>_metrics.txt).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-final}
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -q -m gpu -s
step fuzz 600 python tools/fuzz_gpu_vs_oracle.py --n 400 --seed 33
step bench 900 python bench.py
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
step bench_contracted 900 python bench.py --arithmetic contracted --no-cpu-baseline
tail -n 1 "$OUT/bench_contracted.log" > "$OUT/bench_contracted.json"
NCU="ncu --set full --clock-control none --import-source on"
step ncu_c1 300 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C1" python tools/bench_configs.py --configs C1 --steps 3 --padded 32
step ncu_c2 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C2" python tools/bench_configs.py --configs C2 --steps 3 --padded 32
step ncu_c3a 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C3a" python tools/bench_configs.py --configs C3a --steps 3 --padded 32
step ncu_c3b 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C3b" python tools/bench_configs.py --configs C3b --steps 3 --padded 32
step ncu_c4 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C4" python tools/bench_configs.py --configs C4 --steps 3 --padded 32
step ncu_c5a 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C5A" python tools/bench_configs.py --configs C5 --steps 3 --padded 32
step ncu_c5b 500 $NCU -k regex:k_render -s 39 -c 1 -o "$OUT/prof_C5B" python tools/bench_configs.py --configs C5 --steps 3 --padded 32
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_bench.csv" python bench.py --steps 2 --warmup 1 --configs "" --no-cpu-baseline --e2e-steps 1
export EU_PROFILE_DIR="$OUT/profiles"
step summarise 300 python tools/summarise_kernels.py r02 C1="$OUT/prof_C1.ncu-rep" C2="$OUT/prof_C2.ncu-rep" C3a="$OUT/prof_C3a.ncu-rep" C3b="$OUT/prof_C3b.ncu-rep" C4="$OUT/prof_C4.ncu-rep" C5A="$OUT/prof_C5A.ncu-rep" C5B="$OUT/prof_C5B.ncu-rep"
rm -f "$OUT"/prof_C1.ncu-rep "$OUT"/prof_C3a.ncu-rep "$OUT"/prof_C3b.ncu-rep "$OUT"/prof_C5A.ncu-rep "$OUT"/prof_C5B.ncu-rep
cat "$OUT/summary.txt"
tail -n 3 "$OUT/pytest_gpu.log"; tail -n 2 "$OUT/fuzz.log"
