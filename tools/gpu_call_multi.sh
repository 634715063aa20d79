#!/bin/bash
# One box, N GPUs (gpurun --gpus N): the configs[4] row-band bench on 1, 2 .. N ranks, the reference arm, the replica line.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/multi_n$N
mkdir -p "$OUT"
step() { name=$1; shift; echo "== $name" | tee -a "$OUT/summary.txt"; s=$(date +%s); timeout "$1" "${@:2}" > "$OUT/$name.log" 2>&1; echo "   rc=$? $(( $(date +%s) - s )) s" | tee -a "$OUT/summary.txt"; }
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
lscpu | grep -i -E "numa|model name|^cpu\(s\)" > "$OUT/numa.txt" 2>&1
for n in $N 4 2; do
  if [ $n -le $N ]; then
    step bench_n$n 900 $RUN --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5
    grep '^{"metric"' "$OUT/bench_n$n.log" | tail -n 1 > "$OUT/bench_n$n.json"
  fi
done
step bench_n1_c5 600 python bench.py --workload c5 --steps 20 --warmup 5
grep '^{"metric"' "$OUT/bench_n1_c5.log" | tail -n 1 > "$OUT/bench_n1_c5.json"
if [ "${REPLICAS:-0}" = 1 ]; then
  step bench_n${N}_replicas 900 $RUN --nproc-per-node $N --master-port 29660 bench.py --gpus $N --steps 20 --warmup 5 --workload c2 --partition frames
  grep '^{"metric"' "$OUT/bench_n${N}_replicas.log" | tail -n 1 > "$OUT/bench_n${N}_replicas.json"
fi
step pytest_2gpu 600 python -m pytest tests/test_gpu_peer_frame.py -q -m gpu -s
tail -n 3 "$OUT/pytest_2gpu.log"
cat "$OUT/summary.txt"
