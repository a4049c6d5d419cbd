#!/usr/bin/env bash
# Round-2 multi-GPU pass (gpurun --gpus 8): default build at 8 ranks (BASE caption config) and the LARGE pretrain workload (BASELINE config 5).
set -u
OUT=gpurun_out/r2c11
mkdir -p $OUT
K="timeout -s KILL"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
$K 420 $T bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; head -c 400 $OUT/bench_n8.json; echo; tail -2 $OUT/bench_n8.err
$K 600 $T bench.py --gpus 8 --config large_pretrain224 --steps 10 --warmup 3 > $OUT/bench_n8_large_pretrain224.json 2> $OUT/bench_n8_large.err; head -c 400 $OUT/bench_n8_large_pretrain224.json; echo; tail -2 $OUT/bench_n8_large.err
