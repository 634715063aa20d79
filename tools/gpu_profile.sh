#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench first, then the ncu launch list and one
# full capture of the render kernel of the SAME command. Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BENCH_FLAGS:-}"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 3 -c 1 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/ncu_full_$TAG.log
