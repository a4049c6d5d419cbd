#!/usr/bin/env bash
# Round-2 fifth hardware pass: plain PDL default (GEMM prefetch dropped), parallel fused-LN tail in the decode GEMM, LARGE VQA 480 px,
# gradient parity BASE + LARGE, experiment files removed.
set -u
OUT=gpurun_out/r2c5
mkdir -p $OUT
K="timeout -s KILL"
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
grep -E "grads|bench config|passed|failed|rc=|Error|vs oracle" $OUT/pytest_gpu.log | tail -30
B="$K 600 python bench.py --steps 20 --warmup 5"
PRISMER_BENCH_DUMP=1 $B > $OUT/bench_default.json 2> $OUT/bench_default.err; cp gpurun_out/gemm_shapes.txt $OUT/gemm_shapes.txt
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sec", d.get("secondary",{}).get("value"), d.get("secondary",{}).get("ms_per_step"), (d.get("secondary",{}).get("roofline") or {}).get("frac"), (d.get("secondary",{}).get("roofline") or {}).get("decode_ms_per_batch"))
PY
tail -3 $OUT/bench_default.err
PRISMER_LIB=$PWD/prismer_b200/libprismer_sm100_nopdl.so $B --no-cpu-baseline > $OUT/bench_nopdl.json 2> $OUT/bench_nopdl.err; head -c 200 $OUT/bench_nopdl.json; echo
$K 900 python bench.py --config large_vqa480 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/bench_large_vqa480.json 2> $OUT/bench_large_vqa480.err; head -c 300 $OUT/bench_large_vqa480.json; echo; tail -2 $OUT/bench_large_vqa480.err
$K 900 python bench.py --config base_caption480 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/bench_base_caption480.json 2> $OUT/bench_base_caption480.err; head -c 300 $OUT/bench_base_caption480.json; echo; tail -2 $OUT/bench_base_caption480.err
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; tail -5 $OUT/hbm_kernels.txt
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$K 300 $NCU --metrics $MET -k regex:"skinny|decode_attn" --csv --log-file $OUT/ncu_decode.csv python tools/hbm_kernels.py > $OUT/ncu_decode.log 2>&1
du -sh $OUT
