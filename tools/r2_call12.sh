#!/usr/bin/env bash
# Round-2 hardware pass 12 (1 GPU): rewritten stem plumbing kernels (im2col / BN+ReLU backward) and single-pass CE.
set -u
OUT=gpurun_out/r2c12
mkdir -p $OUT
K="timeout -s KILL"
$K 400 python -m pytest tests/test_bn_kernels_gpu.py tests/test_stem_gpu.py tests/test_loss_gpu.py tests/test_surface_gpu.py tests/test_model_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -6 $OUT/pytest_new.log
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; grep -E "im2col|bn_relu|ce_loss" $OUT/hbm_kernels.txt
B="$K 600 python bench.py --steps 20 --warmup 5"
PRISMER_BENCH_DUMP=1 $B > $OUT/bench_default.json 2> $OUT/bench_default.err; cp gpurun_out/gemm_shapes.txt $OUT/gemm_shapes.txt 2>/dev/null
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sec", d.get("secondary",{}).get("value"), "roof", d["roofline"]["frac"], "loss", d.get("loss"))
PY
tail -3 $OUT/bench_default.err
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
NCU="ncu --clock-control none"
$K 600 $NCU --metrics gpu__time_duration.sum -s 1250 -c 1400 --csv --log-file $OUT/launches_r2b.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/launches_r2b.log 2>&1
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$K 300 $NCU --metrics $MET -k regex:"im2col_nhwc|bn_relu_bwd|bn_bwd_apply|ce_fwd|ce_bwd" --csv --log-file $OUT/ncu_stem.csv python tools/hbm_kernels.py > $OUT/ncu_stem.log 2>&1
du -sh $OUT
