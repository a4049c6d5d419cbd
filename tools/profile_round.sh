#!/usr/bin/env bash
# One-GPU profiling pass for a round (run under gpurun; writes to gpurun_out/, copy the summaries you want judged to profiles/):
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2'
# 1. launch list of one eager training step (per-kernel durations; shares of the step)
# 2. --set full captures of the top GEMM shape and the ViT attention kernels (isolated one-op scripts: few replays)
# Never run the step under torch.profiler / CUPTI here -- it hung a box in round 1.
set -u
R=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -s 1300 -c 1400 --csv --log-file $OUT/launches_$R.csv \
    python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_$R.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -o $OUT/gemm_$R -f \
    python tools/one_gemm.py 8320 3072 768 > $OUT/gemm_$R.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:attn_ -s 9 -c 3 -o $OUT/attn_$R -f \
    python tools/one_attn.py 32 12 260 260 64 > $OUT/attn_$R.log 2>&1
python - <<PY
import csv, collections
rows = [l for l in open("$OUT/launches_$R.csv") if l.startswith('"')]
rd = csv.reader(rows); hdr = next(rd); ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    try:
        k = r[ki].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", ""); t = float(r[vi].replace(",", "")) / 1e3
    except Exception:
        continue
    agg[k][0] += 1; agg[k][1] += t
tot = sum(v[1] for v in agg.values())
with open("$OUT/launches_${R}_summary.txt", "w") as f:
    f.write(f"# {sum(v[0] for v in agg.values())} launches, total {tot / 1e3:.2f} ms (serialized, cold cache: compare shares)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  {k}\n")
print(open("$OUT/launches_${R}_summary.txt").read()[:3000])
PY
