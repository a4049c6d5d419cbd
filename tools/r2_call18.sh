#!/usr/bin/env bash
# Round-2 hardware pass 18 (1 GPU): LARGE / 480-px workloads on the final build.
set -u
OUT=gpurun_out/r2c18
mkdir -p $OUT
K="timeout -s KILL"
F="--no-secondary --no-cpu-baseline"
for c in large_pretrain224 large_vqa480 base_caption480; do
$K 900 python bench.py --config $c --steps 10 --warmup 3 $F > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "$c: $(head -c 330 $OUT/bench_$c.json)"; tail -1 $OUT/bench_$c.err
done
du -sh $OUT
