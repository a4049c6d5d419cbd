#!/usr/bin/env bash
# Round-2 multi-GPU pass (gpurun --gpus 8) on the final build: BASE caption step and LARGE pretrain step (BASELINE config 5) at 8 ranks.
set -u
OUT=gpurun_out/r2c20
mkdir -p $OUT
K="timeout -s KILL"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
F="--no-secondary --no-cpu-baseline"
$K 300 $T bench.py --gpus 8 --config large_pretrain224 --steps 10 --warmup 3 $F > $OUT/bench_n8_large_pretrain224.json 2> $OUT/bench_n8_large.err; echo "large: $(grep '^{' $OUT/bench_n8_large_pretrain224.json | head -c 300)"; tail -2 $OUT/bench_n8_large.err
$K 240 $T bench.py --gpus 8 --steps 20 --warmup 5 $F > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "base: $(grep '^{' $OUT/bench_n8.json | head -c 300)"; tail -2 $OUT/bench_n8.err
