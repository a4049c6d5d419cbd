#!/usr/bin/env bash
# Round-2 multi-GPU pass (gpurun --gpus 2): the default PDL build under NCCL, and the all-reduces captured inside the step's CUDA graph.
set -u
OUT=gpurun_out/r2c6
mkdir -p $OUT
K="timeout -s KILL"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$K 600 $T bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_default.json 2> $OUT/bench_n2_default.err; head -c 300 $OUT/bench_n2_default.json; echo; tail -3 $OUT/bench_n2_default.err
$K 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --comm-in-graph > $OUT/bench_n2_ingraph.json 2> $OUT/bench_n2_ingraph.err; head -c 300 $OUT/bench_n2_ingraph.json; echo; tail -5 $OUT/bench_n2_ingraph.err
$K 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; head -c 200 $OUT/bench_n1.json; echo
