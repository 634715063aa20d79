#!/bin/bash
# One box, one GPU: the whole GPU suite, sweeps, bench lines, kernel variants, one ncu capture per config kernel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-call3}
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }

step pytest_gpu 1500 python -m pytest tests -q -m gpu -s
step fuzz_gpu 400 python tools/fuzz_gpu_vs_oracle.py --n 300 --seed 21
step fuzz_gpu_contracted 400 python tools/fuzz_gpu_vs_oracle.py --n 200 --seed 22 --contracted 1
step bench 900 python bench.py --arithmetic exact
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
step bench_contracted 900 python bench.py --arithmetic contracted --no-cpu-baseline
tail -n 1 "$OUT/bench_contracted.log" > "$OUT/bench_contracted.json"
step configs_variants 600 python tools/bench_configs.py --configs C1,C2,C3a,C3b,C4 --padded 0,8,1,9 --steps 10
step bench_c5_n1 600 python bench.py --workload c5 --steps 10 --warmup 3
tail -n 1 "$OUT/bench_c5_n1.log" > "$OUT/bench_c5_n1.json"
step reference_c5 900 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_bench.csv" \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs ""
NCU="ncu --set full --clock-control none --import-source on"
step ncu_c2 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C2" python tools/bench_configs.py --configs C2 --steps 3
step ncu_c2_contracted 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C2_contracted" python tools/bench_configs.py --configs C2 --steps 3 --padded 8
step ncu_c1 300 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C1" python tools/bench_configs.py --configs C1 --steps 3
step ncu_c3a 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C3a" python tools/bench_configs.py --configs C3a --steps 3
step ncu_c3b 600 $NCU -k regex:k_render -s 9 -c 1 -o "$OUT/prof_C3b" python tools/bench_configs.py --configs C3a,C3b --steps 3
step ncu_c4 400 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C4" python tools/bench_configs.py --configs C4 --steps 3
step ncu_c5a 500 $NCU -k regex:k_render -s 3 -c 1 -o "$OUT/prof_C5A" python tools/bench_configs.py --configs C5 --steps 3
step ncu_c5b 500 $NCU -k regex:k_render -s 39 -c 1 -o "$OUT/prof_C5B" python tools/bench_configs.py --configs C5 --steps 3
step ncu_stage 400 $NCU -k regex:"k_iir|k_cm_" -c 12 -o "$OUT/prof_stage" python tools/bench_configs.py --configs C2 --steps 1
export EU_PROFILE_DIR="$OUT/profiles"
step summarise 300 python tools/summarise_kernels.py r02 C1="$OUT/prof_C1.ncu-rep" C2="$OUT/prof_C2.ncu-rep" C2/contracted="$OUT/prof_C2_contracted.ncu-rep" \
  C3a="$OUT/prof_C3a.ncu-rep" C3b="$OUT/prof_C3b.ncu-rep" C4="$OUT/prof_C4.ncu-rep" C5A="$OUT/prof_C5A.ncu-rep" C5B="$OUT/prof_C5B.ncu-rep"
ncu -i "$OUT/prof_stage.ncu-rep" --page raw --csv > "$OUT/stage_raw.csv" 2>/dev/null
rm -f "$OUT"/prof_C1.ncu-rep "$OUT"/prof_C3a.ncu-rep "$OUT"/prof_C3b.ncu-rep "$OUT"/prof_C5A.ncu-rep "$OUT"/prof_C5B.ncu-rep \
  "$OUT"/prof_C2_contracted.ncu-rep "$OUT"/prof_stage.ncu-rep
du -sh "$OUT" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
