#!/usr/bin/env bash
# Round-2 fourth hardware pass: PDL is the default build (with L2 prefetch of weight tiles before the dependency wait), decode kernels
# rewritten for memory-level parallelism, LARGE workloads, gradient parity for BASE + LARGE.
set -u
OUT=gpurun_out/r2c4
mkdir -p $OUT
K="timeout -s KILL"
$K 300 python -m pytest tests/test_kv_decode_gpu.py tests/test_gemm_gpu.py tests/test_layernorm_gpu.py -q -x -p no:cacheprovider -s > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -6 $OUT/pytest_new.log
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
PRISMER_LIB=$PWD/prismer_b200/libprismer_sm100_nopdl.so $B --no-cpu-baseline --no-secondary > $OUT/bench_nopdl.json 2> $OUT/bench_nopdl.err; head -c 200 $OUT/bench_nopdl.json; echo
$K 600 python -c "
import sys; sys.path.insert(0,'.')
import torch
from prismer_b200 import generation
import bench, types
# cache-less schedule (the reference's): captions/s for comparison with the KV-cached default
generation.KV_CACHE = False
sys.argv=['bench.py','--mode','caption','--steps','10','--warmup','3']
bench.main()
" > $OUT/bench_caption_nocache.json 2> $OUT/bench_caption_nocache.err; head -c 250 $OUT/bench_caption_nocache.json; echo
$K 900 python bench.py --config large_pretrain224 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/bench_large_pretrain224.json 2> $OUT/bench_large_pretrain224.err; head -c 300 $OUT/bench_large_pretrain224.json; echo; tail -2 $OUT/bench_large_pretrain224.err
$K 900 python bench.py --config large_vqa480 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/bench_large_vqa480.json 2> $OUT/bench_large_vqa480.err; head -c 300 $OUT/bench_large_vqa480.json; echo; tail -2 $OUT/bench_large_vqa480.err
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; tail -5 $OUT/hbm_kernels.txt
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$K 300 $NCU --metrics $MET -k regex:"skinny|decode_attn" --csv --log-file $OUT/ncu_decode.csv python tools/hbm_kernels.py > $OUT/ncu_decode.log 2>&1
$K 600 $NCU --metrics gpu__time_duration.sum -s 1250 -c 1400 --csv --log-file $OUT/launches_r2.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/launches_r2.log 2>&1
du -sh $OUT
