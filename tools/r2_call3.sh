#!/usr/bin/env bash
# Round-2 third hardware pass: 8-warp attention epilogues, KV-cached decode, new parity tests, PDL build, bench with secondary metrics,
# and SMALL ncu outputs (metrics-only CSVs + two single-kernel full captures; < 64 MiB in total).
set -u
OUT=gpurun_out/r2c3
mkdir -p $OUT
K="timeout -s KILL"
$K 300 python -m pytest tests/test_attention_gpu.py tests/test_kv_decode_gpu.py -q -x -p no:cacheprovider -s > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
tail -8 $OUT/pytest_new.log
$K 120 python tools/one_attn.py 32 12 260 260 64 > $OUT/one_attn.log 2>&1; tail -1 $OUT/one_attn.log
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
grep -E "BASE|bench config|passed|failed|rc=|asserted|ids asserted" $OUT/pytest_gpu.log | tail -30
B="$K 600 python bench.py --steps 20 --warmup 5"
PRISMER_BENCH_DUMP=1 $B > $OUT/bench_default.json 2> $OUT/bench_default.err; cp gpurun_out/gemm_shapes.txt $OUT/gemm_shapes.txt
head -c 300 $OUT/bench_default.json; echo; tail -3 $OUT/bench_default.err
$K 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; head -c 400 $OUT/bench_reference.json; echo
export PRISMER_LIB=$PWD/prismer_b200/libprismer_sm100_pdl.so
$K 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_pdl.log 2>&1; echo "rc=$?" >> $OUT/pytest_pdl.log
tail -3 $OUT/pytest_pdl.log
$B --no-cpu-baseline > $OUT/bench_pdl.json 2> $OUT/bench_pdl.err; head -c 300 $OUT/bench_pdl.json; echo
unset PRISMER_LIB
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
$K 300 $NCU --metrics $MET --csv --log-file $OUT/ncu_hbm.csv python tools/hbm_kernels.py > $OUT/ncu_hbm.log 2>&1
$K 200 $NCU --metrics $MET,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:gemm_bf16_kernel -s 3 -c 1 --csv --log-file $OUT/ncu_gemm_big.csv python tools/one_gemm.py 8320 3072 768 > /dev/null 2>&1
$K 200 $NCU --set full --import-source on -k regex:attn_ -s 6 -c 2 -o $OUT/attn_tc_r2 -f python tools/one_attn.py 32 12 260 260 64 > $OUT/ncu_attn.log 2>&1
$K 200 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -o $OUT/gemm_small_r2 -f python tools/one_gemm.py 960 768 768 > $OUT/ncu_gemm_small.log 2>&1
$K 600 $NCU --metrics gpu__time_duration.sum -s 1250 -c 1400 --csv --log-file $OUT/launches_r2.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/launches_r2.log 2>&1
du -sh $OUT; ls -la $OUT | head -40
