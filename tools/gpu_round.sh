#!/bin/bash
# One GPU visit: GPU tests, bench line, per-launch time + DRAM-byte lists of one step, full ncu capture of the hot kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
KRE='regex:^(attn|gemm|rmsnorm|swiglu|build_h0|ce_|qav|sum_scale|f32_to|video|visual|scatter|option)'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt
if [ "${2:-}" != "skip_tests" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json
BENCH1="python bench.py --steps 1 --warmup 3 --no-e2e --no-padfree --no-cpu-baseline --sample-layers 0"
$BENCH1 > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -s 1250 -c 620 --csv \
    --log-file $OUT/launches_$TAG.csv $BENCH1 > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"
python tools/attn_bench.py > $OUT/attn_bench_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_nt|attn_|rmsnorm_bwd" -s 24 -c 6 \
    -f -o $OUT/prof_$TAG python tools/attn_bench.py > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
cat $OUT/attn_bench_$TAG.log
