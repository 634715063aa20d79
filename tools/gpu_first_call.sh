#!/bin/bash
# Everything that was written after the GPU time of round 1 had run out, in the order of what it can break:
# run as ONE gpurun call at the start of the next round,
#   gpurun --timeout 2700 -- 'bash tools/gpu_first_call.sh'
# and read gpurun_out/first_call/. Each step has its own time limit and its own process, so that a fault on
# an untried size cannot take the later steps with it.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/first_call
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$?" | tee -a "$OUT/summary.txt"; }

# 1. the must-pass set, as the driver runs it
step pytest_gpu 900 python -m pytest tests -x -q -m gpu
# 2. the opt-in tests: degenerate sizes, and the FMA-window build against the contracted oracle
EU_GPU_UNTRIED=1 step pytest_untried 600 python -m pytest tests/test_gpu_parity.py::test_edge_jobs tests/test_contracted.py -q -m gpu -rxXs
# 3. random jobs, kernels vs oracle (tools/fuzz_oracle_vs_reference.py's generator)
step fuzz_gpu 600 python tools/fuzz_gpu_vs_oracle.py --n 500
# 4. the bench in both arithmetics, back to back on the same box (the default first: it is the headline)
step bench_exact 600 python bench.py
EU_ARITHMETIC=contracted step bench_contracted 600 python bench.py --no-cpu-baseline
tail -n 1 "$OUT/bench_exact.log" > "$OUT/bench_exact.json"
tail -n 1 "$OUT/bench_contracted.log" > "$OUT/bench_contracted.json"
# 5. launch list of the contracted build (share of the step per kernel), only after its plain run ended well
if grep -q '"value"' "$OUT/bench_contracted.json"; then
  EU_ARITHMETIC=contracted step ncu_launches_contracted 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/launches_contracted.csv" python bench.py --steps 20 --warmup 3 --no-cpu-baseline
fi
# 6. LAST, because a mistake in it would hang rather than fail: the per-warp staging kernel (k_render_warp).
#    Parity first (own process, four-minute limit inside the test), the bench only if that passed.
EU_GPU_UNTRIED=1 step pytest_warp 330 python -m pytest tests/test_gpu_parity.py::test_warp_staged_kernel_is_bit_exact -q -rxXs
if grep -q "1 passed" "$OUT/pytest_warp.log"; then
  step fuzz_gpu_warp 400 python tools/fuzz_gpu_vs_oracle.py --n 300 --seed 12 --warp-tiles 1
  step bench_warp 300 python bench.py --warp-tiles 1 --no-cpu-baseline
  tail -n 1 "$OUT/bench_warp.log" > "$OUT/bench_warp.json"
  EU_ARITHMETIC=contracted step bench_warp_contracted 300 python bench.py --warp-tiles 1 --no-cpu-baseline
  tail -n 1 "$OUT/bench_warp_contracted.log" > "$OUT/bench_warp_contracted.json"
  # launch list + one full capture of the per-warp kernel (tools/gpu_profile.sh writes gpurun_out/*_r02warp*)
  BENCH_FLAGS="--warp-tiles 1" step profile_warp 420 bash tools/gpu_profile.sh r02warp
  # every single-facet config, block-staged / direct (0) against per-warp staging (4), same inputs
  step configs_warp 420 python tools/bench_configs.py --configs C1,C2,C3a,C3b,C4 --padded 0,4 --steps 10
fi
cat "$OUT/summary.txt"
