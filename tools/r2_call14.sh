#!/usr/bin/env bash
# Round-2 hardware pass 14 (1 GPU): instance-embedding gradient through a shared-memory slab, bn_stats / im2col_first on the new thread layout.
set -u
OUT=gpurun_out/r2c14
mkdir -p $OUT
K="timeout -s KILL"
$K 400 python -m pytest tests/test_bn_kernels_gpu.py tests/test_stem_gpu.py tests/test_loss_gpu.py tests/test_model_gpu.py tests/test_zz_base_grads_gpu.py tests/test_zzzz_compact_gpu.py tests/test_surface_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -4 $OUT/pytest_new.log
B="$K 600 python bench.py --steps 20 --warmup 5"
$B > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sec", d.get("secondary",{}).get("value"), d.get("secondary",{}).get("ms_per_step"), "loss", d.get("loss"))
ep=d.get("entry_point_ms_per_step",{})
print({k:ep[k] for k in ("bn_stats","im2col_first","assemble_tokens_bwd","bn_relu_bwd","im2col_nhwc") if k in ep})
PY
tail -3 $OUT/bench_default.err
du -sh $OUT
