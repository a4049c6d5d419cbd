#!/usr/bin/env bash
# Round-2 sixth hardware pass (1 GPU): decode kernels v3 (register-resident LN tail, adaptive K chunk) + concurrent decode chains.
set -u
OUT=gpurun_out/r2c6
mkdir -p $OUT
K="timeout -s KILL"
$K 300 python -m pytest tests/test_kv_decode_gpu.py tests/test_model_gpu.py tests/test_zzz_beam_gpu.py -q -x -p no:cacheprovider -s > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -8 $OUT/pytest_new.log
$K 600 python bench.py --mode caption --steps 20 --warmup 5 > $OUT/bench_caption.json 2> $OUT/bench_caption.err; head -c 400 $OUT/bench_caption.json; echo; tail -3 $OUT/bench_caption.err
for n in 1 2 8; do
$K 300 python -c "
import sys; sys.path.insert(0,'.')
from prismer_b200 import kv_decode
kv_decode.DECODE_CHAINS = $n
import bench
sys.argv=['bench.py','--mode','caption','--steps','10','--warmup','3']
bench.main()
" > $OUT/bench_caption_chains$n.json 2> $OUT/bench_caption_chains$n.err; echo "chains=$n: $(head -c 200 $OUT/bench_caption_chains$n.json)"
done
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; tail -5 $OUT/hbm_kernels.txt
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
