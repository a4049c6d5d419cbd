#!/usr/bin/env bash
# Round-2 multi-GPU pass (gpurun --gpus 4): BASE caption step and LARGE pretrain step at N = 4 on the final build.
set -u
OUT=gpurun_out/r2c19
mkdir -p $OUT
K="timeout -s KILL"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
F="--no-secondary --no-cpu-baseline"
$K 600 $T bench.py --gpus 4 --steps 20 --warmup 5 $F > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo "base: $(grep '^{' $OUT/bench_n4.json | head -c 260)"; tail -2 $OUT/bench_n4.err
$K 900 $T bench.py --gpus 4 --config large_pretrain224 --steps 10 --warmup 3 $F > $OUT/bench_n4_large_pretrain224.json 2> $OUT/bench_n4_large.err; echo "large: $(grep '^{' $OUT/bench_n4_large_pretrain224.json | head -c 300)"; tail -2 $OUT/bench_n4_large.err
du -sh $OUT
