#!/usr/bin/env bash
# Round-2 hardware pass 22 (1 GPU): the surface tests after their three-segment expectation was updated.
set -u
OUT=gpurun_out/r2c22
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests/test_surface_gpu.py -q -p no:cacheprovider > $OUT/pytest_surface.log 2>&1; echo "rc=$?" >> $OUT/pytest_surface.log
tail -4 $OUT/pytest_surface.log
