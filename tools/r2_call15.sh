#!/usr/bin/env bash
# Round-2 hardware pass 15 (1 GPU): KV-cache prefill of the prompt in one full-sequence pass; A/B against token-by-token.
set -u
OUT=gpurun_out/r2c15
mkdir -p $OUT
K="timeout -s KILL"
$K 400 python -m pytest tests/test_kv_decode_gpu.py tests/test_zzz_beam_gpu.py tests/test_zzz_surface_golden_gpu.py tests/test_zzz_properties_gpu.py tests/test_model_gpu.py tests/test_surface_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -4 $OUT/pytest_new.log
$K 300 python bench.py --mode caption --steps 10 --warmup 3 > $OUT/bench_caption.json 2> $OUT/bench_caption.err; echo "prefill: $(head -c 260 $OUT/bench_caption.json)"; tail -2 $OUT/bench_caption.err
$K 300 python -c "
import sys; sys.path.insert(0,'.')
from prismer_b200 import kv_decode
kv_decode.PREFILL = False
import bench
sys.argv=['bench.py','--mode','caption','--steps','10','--warmup','3']
bench.main()
" > $OUT/bench_caption_noprefill.json 2> $OUT/bench_caption_noprefill.err; echo "token-by-token: $(head -c 260 $OUT/bench_caption_noprefill.json)"
du -sh $OUT
