#!/usr/bin/env bash
# Round-2 hardware pass 21 (1 GPU): final tree -- full GPU suite, smoke, default bench line; 64-column LM-head decode GEMM A/B.
set -u
OUT=gpurun_out/r2c21
mkdir -p $OUT
K="timeout -s KILL"
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
$K 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
$K 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
s=d.get("secondary",{})
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "mfu", d["step_mfu"]["frac_of_peak"], "roof", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
print("caption", s.get("value"), s.get("ms_per_step"), (s.get("roofline") or {}).get("frac"), (s.get("roofline") or {}).get("decode_ms_per_batch"))
PY
tail -2 $OUT/bench_default.err
PRISMER_SKINNY_WIDE=0 $K 300 python bench.py --mode caption --steps 10 --warmup 3 > $OUT/bench_caption_narrow.json 2> $OUT/bench_caption_narrow.err; echo "narrow LM head: $(head -c 200 $OUT/bench_caption_narrow.json)"
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; grep -E "skinny|decode_att" $OUT/hbm_kernels.txt
du -sh $OUT
