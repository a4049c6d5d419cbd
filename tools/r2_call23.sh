#!/usr/bin/env bash
# Round-2 hardware pass 23 (1 GPU): ncu launch list of one eager training step on the final tree (shares of the step per kernel).
set -u
OUT=gpurun_out/r2c23
mkdir -p $OUT
K="timeout -s KILL"
NCU="ncu --clock-control none"
$K 400 $NCU --metrics gpu__time_duration.sum -s 1250 -c 1400 --csv --log-file $OUT/launches_r2c.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/launches_r2c.log 2>&1
tail -2 $OUT/launches_r2c.log
du -sh $OUT
