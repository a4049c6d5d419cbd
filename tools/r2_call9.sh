#!/usr/bin/env bash
# Round-2 seventh hardware pass (1 GPU): fused-LN decode GEMM without MEMBAR.SC; decode chains 1 / 2 / 4 / 8.
set -u
OUT=gpurun_out/r2c9
mkdir -p $OUT
K="timeout -s KILL"
$K 300 python -m pytest tests/test_kv_decode_gpu.py tests/test_model_gpu.py tests/test_zzz_beam_gpu.py tests/test_zzz_surface_golden_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -4 $OUT/pytest_new.log
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; tail -5 $OUT/hbm_kernels.txt
for n in 1 2 4; do
$K 300 python -c "
import sys; sys.path.insert(0,'.')
from prismer_b200 import kv_decode
kv_decode.DECODE_CHAINS = $n
import bench
sys.argv=['bench.py','--mode','caption','--steps','10','--warmup','3']
bench.main()
" > $OUT/bench_caption_chains$n.json 2> $OUT/bench_caption_chains$n.err; echo "chains=$n: $(head -c 200 $OUT/bench_caption_chains$n.json)"; tail -2 $OUT/bench_caption_chains$n.err
done
PRISMER_LIB=$PWD/prismer_b200/libprismer_sm100_nopdl.so $K 300 python bench.py --mode caption --steps 10 --warmup 3 > $OUT/bench_caption_nopdl.json 2> $OUT/bench_caption_nopdl.err; echo "nopdl: $(head -c 200 $OUT/bench_caption_nopdl.json)"
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$K 300 $NCU --metrics $MET -k regex:"skinny|decode_attn" --csv --log-file $OUT/ncu_decode.csv python tools/hbm_kernels.py > $OUT/ncu_decode.log 2>&1
du -sh $OUT
