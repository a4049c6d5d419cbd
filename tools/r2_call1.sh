#!/usr/bin/env bash
# Round-2 first hardware pass: validate everything that was committed without a GPU at the end of round 1.
set -u
OUT=gpurun_out/r2c1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
timeout 600 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
export PRISMER_EXPERIMENTAL=1
for k in unfused_attention two_cta layernorm_bwd_v2 cross_shapes bn192; do
  timeout 150 python -m pytest tests/test_experimental_gpu.py -q -x -s -k $k -p no:cacheprovider > $OUT/exp_$k.log 2>&1; echo "rc=$?" >> $OUT/exp_$k.log
done
timeout 150 python tools/bench_attn_unfused.py > $OUT/bench_attn_unfused.log 2>&1; echo "rc=$?" >> $OUT/bench_attn_unfused.log
timeout 150 python tools/bench_gemm_2cta.py > $OUT/bench_gemm_2cta.log 2>&1; echo "rc=$?" >> $OUT/bench_gemm_2cta.log
unset PRISMER_EXPERIMENTAL
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
PRISMER_BENCH_DUMP=1 $B > $OUT/bench_default.json 2> $OUT/bench_default.err; cp gpurun_out/gemm_shapes.txt $OUT/gemm_shapes_default.txt
PRISMER_ATTN_UNFUSED=bwd $B > $OUT/bench_attn_unfused_bwd.json 2> $OUT/bench_attn_unfused_bwd.err
PRISMER_ATTN_UNFUSED=1 $B > $OUT/bench_attn_unfused_all.json 2> $OUT/bench_attn_unfused_all.err
PRISMER_LN_BWD_V2=1 $B > $OUT/bench_lnv2.json 2> $OUT/bench_lnv2.err
PRISMER_GEMM_BN192=1 $B > $OUT/bench_bn192.json 2> $OUT/bench_bn192.err
PRISMER_GEMM_2CTA=1 $B > $OUT/bench_2cta.json 2> $OUT/bench_2cta.err
$B --compact-inputs > $OUT/bench_compact.json 2> $OUT/bench_compact.err
$B --overlap-optimizer > $OUT/bench_overlapopt.json 2> $OUT/bench_overlapopt.err
$B --mode caption > $OUT/bench_caption.json 2> $OUT/bench_caption.err
tail -c 600 $OUT/pytest_gpu.log
for f in $OUT/bench_*.json; do echo "$f: $(head -c 200 $f)"; done
