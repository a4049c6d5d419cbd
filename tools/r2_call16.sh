#!/usr/bin/env bash
# Round-2 hardware pass 16 (1 GPU): full GPU suite, smoke, default bench line and the reference arm on the final tree.
set -u
OUT=gpurun_out/r2c16
mkdir -p $OUT
K="timeout -s KILL"
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
$K 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
B="$K 600 python bench.py --steps 20 --warmup 5"
PRISMER_BENCH_DUMP=1 $B > $OUT/bench_default.json 2> $OUT/bench_default.err; cp gpurun_out/gemm_shapes.txt $OUT/gemm_shapes.txt 2>/dev/null
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
s=d.get("secondary",{})
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "mfu", d["step_mfu"]["frac_of_peak"], "roof", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
print("caption", s.get("value"), s.get("ms_per_step"), (s.get("roofline") or {}).get("frac"), (s.get("roofline") or {}).get("decode_ms_per_batch"), (s.get("roofline") or {}).get("encoder_ms_per_batch"))
PY
tail -2 $OUT/bench_default.err
$K 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; head -c 300 $OUT/bench_reference.json; echo
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1
du -sh $OUT
