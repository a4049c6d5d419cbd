#!/usr/bin/env bash
# Round-2 hardware pass 10 (1 GPU): decode with LayerNorm-on-load, rank without tile, full suite + default bench.
set -u
OUT=gpurun_out/r2c10
mkdir -p $OUT
K="timeout -s KILL"
$K 300 python -m pytest tests/test_kv_decode_gpu.py tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_zzz_beam_gpu.py tests/test_zzz_surface_golden_gpu.py tests/test_surface_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -4 $OUT/pytest_new.log
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; tail -5 $OUT/hbm_kernels.txt
$K 300 python bench.py --mode caption --steps 20 --warmup 5 > $OUT/bench_caption.json 2> $OUT/bench_caption.err; echo "caption: $(head -c 220 $OUT/bench_caption.json)"; tail -2 $OUT/bench_caption.err
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
$K 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sec", d.get("secondary",{}).get("value"), d.get("secondary",{}).get("ms_per_step"), (d.get("secondary",{}).get("roofline") or {}).get("frac"), (d.get("secondary",{}).get("roofline") or {}).get("decode_ms_per_batch"))
PY
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$K 300 $NCU --metrics $MET -k regex:"skinny|decode_attn" --csv --log-file $OUT/ncu_decode.csv python tools/hbm_kernels.py > $OUT/ncu_decode.log 2>&1
du -sh $OUT
