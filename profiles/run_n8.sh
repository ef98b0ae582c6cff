#!/bin/bash
# The 8-GPU bundle of a round as ONE gpurun call:   gpurun --gpus 8 --timeout 900 -- 'bash profiles/run_n8.sh <tag>'
#   config3 island GA (the scored bench line), config5 island GA, config4 (48 groups farmed over the GPUs).
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== config3, $N GPUs"
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_n$N.json
echo "== config5, $N GPUs"
timeout 400 $TR --master-port 29522 bench.py --workload config5 --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_config5_n$N.json 2> $OUT/${TAG}_bench_config5_n$N.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_config5_n$N.json
echo "== config4: 48 groups over $N GPUs"
timeout 400 python profiles/run_config4.py 48 --cpu 0 > $OUT/${TAG}_config4_48groups_${N}gpu.json 2> $OUT/${TAG}_config4.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_config4_48groups_${N}gpu.json
echo "== done"
