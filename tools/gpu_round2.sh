#!/bin/bash
# One GPU visit (round 2): GPU tests, default bench line, per-launch time + DRAM-byte list of one step under ncu.
# usage (under gpurun): bash tools/gpu_round2.sh <tag> [skip_tests]
set -u
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
KRE='regex:^(attn|gemm|rmsnorm|swiglu|build_h0|ce_|qav|sum_scale|f32_to|video|visual|linear|scatter|option|grad_scale|scale_f32|gather|expand|move)'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt
if [ "${2:-}" != "skip_tests" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json
BENCH1="python bench.py --steps 1 --warmup 3 --no-e2e --no-padfree --no-cpu-baseline --no-eager-baseline --sample-layers 0"
$BENCH1 > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KRE" -s 1300 -c 640 --csv \
    --log-file $OUT/launches_$TAG.csv $BENCH1 > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"
