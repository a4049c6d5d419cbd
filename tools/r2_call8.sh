#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2c8
mkdir -p $OUT
K="timeout -s KILL"
NCU="ncu --clock-control none"
$K 300 $NCU --set full --import-source on -k regex:skinny -s 7 -c 3 -o $OUT/skinny_r2 -f python tools/hbm_kernels.py > $OUT/ncu_skinny.log 2>&1
ls -la $OUT
