#!/usr/bin/env bash
# Round-2 hardware pass 13 (1 GPU): expert stems as parallel graph branches, A/B against the serial order.
set -u
OUT=gpurun_out/r2c13
mkdir -p $OUT
K="timeout -s KILL"
$K 400 python -m pytest tests/test_loss_gpu.py tests/test_stem_gpu.py tests/test_surface_gpu.py tests/test_model_gpu.py tests/test_zz_base_grads_gpu.py tests/test_zzz_surface_golden_gpu.py tests/test_zzzz_compact_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -4 $OUT/pytest_new.log
B="$K 600 python bench.py --steps 20 --warmup 5"
$B > $OUT/bench_branches.json 2> $OUT/bench_branches.err
PRISMER_STEM_BRANCHES=0 $B --no-cpu-baseline > $OUT/bench_serial.json 2> $OUT/bench_serial.err
python - <<PY
import json
for f in ("branches","serial"):
    try:
        d=json.loads(open("$OUT/bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sec", d.get("secondary",{}).get("value"), d.get("secondary",{}).get("ms_per_step"), "loss", d.get("loss"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 $OUT/bench_branches.err
du -sh $OUT
