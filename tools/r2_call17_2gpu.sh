#!/usr/bin/env bash
# Round-2 multi-GPU pass (gpurun --gpus 2): three-segment gradient overlap (stems' slice reduced last) against the two-segment version;
# LARGE pretrain step at N = 2.
set -u
OUT=gpurun_out/r2c17
mkdir -p $OUT
K="timeout -s KILL"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
F="--no-secondary --no-cpu-baseline"
$K 600 $T bench.py --gpus 2 --steps 20 --warmup 5 $F > $OUT/bench_n2_seg3.json 2> $OUT/bench_n2_seg3.err; echo "seg3: $(head -c 260 $OUT/bench_n2_seg3.json)"; tail -2 $OUT/bench_n2_seg3.err
PRISMER_DP_SEGMENTS=2 $K 600 $T bench.py --gpus 2 --steps 20 --warmup 5 $F > $OUT/bench_n2_seg2.json 2> $OUT/bench_n2_seg2.err; echo "seg2: $(head -c 260 $OUT/bench_n2_seg2.json)"; tail -2 $OUT/bench_n2_seg2.err
$K 900 $T bench.py --gpus 2 --config large_pretrain224 --steps 10 --warmup 3 $F > $OUT/bench_n2_large_pretrain224.json 2> $OUT/bench_n2_large.err; echo "large: $(head -c 300 $OUT/bench_n2_large_pretrain224.json)"; tail -2 $OUT/bench_n2_large.err
du -sh $OUT
