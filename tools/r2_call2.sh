#!/usr/bin/env bash
# Round-2 second hardware pass: first run of the tcgen05 / TMEM attention kernels, lean LayerNorm backward as default, PDL build variant,
# ncu evidence (launch list, --set full on the attention kernels, a decoder-sized GEMM and the HBM-bound kernels).
set -u
OUT=gpurun_out/r2c2
mkdir -p $OUT
K="timeout -s KILL"
$K 240 python -m pytest tests/test_attention_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_attn.log 2>&1; echo "rc=$?" >> $OUT/pytest_attn.log
tail -5 $OUT/pytest_attn.log
if ! grep -q "rc=0" $OUT/pytest_attn.log; then
  # which cases fail? run them one by one without -x (bounded)
  $K 300 python -m pytest tests/test_attention_gpu.py -q -p no:cacheprovider > $OUT/pytest_attn_all.log 2>&1; echo "rc=$?" >> $OUT/pytest_attn_all.log
  tail -30 $OUT/pytest_attn_all.log
fi
$K 120 python tools/one_attn.py 32 12 260 260 64 > $OUT/one_attn.log 2>&1; cat $OUT/one_attn.log | tail -2
$K 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
B="$K 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
PRISMER_BENCH_DUMP=1 $B > $OUT/bench_default.json 2> $OUT/bench_default.err; cp gpurun_out/gemm_shapes.txt $OUT/gemm_shapes.txt
$B --compact-inputs > $OUT/bench_compact.json 2> $OUT/bench_compact.err
export PRISMER_LIB=$PWD/prismer_b200/libprismer_sm100_pdl.so
$K 300 python -m pytest tests/test_gemm_gpu.py tests/test_layernorm_gpu.py tests/test_model_gpu.py -q -x -p no:cacheprovider > $OUT/pytest_pdl.log 2>&1; echo "rc=$?" >> $OUT/pytest_pdl.log
tail -3 $OUT/pytest_pdl.log
$B > $OUT/bench_pdl.json 2> $OUT/bench_pdl.err
unset PRISMER_LIB
python tools/hbm_kernels.py > $OUT/hbm_kernels.txt 2>&1; cat $OUT/hbm_kernels.txt
NCU="ncu --clock-control none"
$K 300 $NCU --set full --import-source on -o $OUT/hbm_r2 -f python tools/hbm_kernels.py > $OUT/ncu_hbm.log 2>&1
$K 200 $NCU --set full --import-source on -k regex:attn_ -s 6 -c 2 -o $OUT/attn_tc_r2 -f python tools/one_attn.py 32 12 260 260 64 > $OUT/ncu_attn.log 2>&1
$K 200 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -o $OUT/gemm_small_r2 -f python tools/one_gemm.py 960 768 768 > $OUT/ncu_gemm_small.log 2>&1
$K 200 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -o $OUT/gemm_mid_r2 -f python tools/one_gemm.py 8320 768 768 > $OUT/ncu_gemm_mid.log 2>&1
$K 600 $NCU --metrics gpu__time_duration.sum -s 1250 -c 1400 --csv --log-file $OUT/launches_r2.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_r2.log 2>&1
for f in $OUT/bench_*.json; do echo "$f: $(head -c 160 $f)"; done
ls -la $OUT
