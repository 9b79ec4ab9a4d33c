#!/bin/bash
# 8-GPU visit: default bench line + two exchange variants (one box, back to back). usage under gpurun --gpus 8: bash tools/scale8.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-padfree"
$RUN --steps 20 --warmup 5 > $OUT/${TAG}_bench_8gpu.json 2> $OUT/${TAG}_bench_8gpu.err; echo "default exit $?"
NCCL_MAX_CTAS=4 $RUN --steps 20 --warmup 5 --no-e2e > $OUT/${TAG}_bench_8gpu_maxctas4.json 2> $OUT/${TAG}_bench_8gpu_maxctas4.err; echo "maxctas exit $?"
$RUN --steps 20 --warmup 5 --no-e2e --chunk-layers 32 > $OUT/${TAG}_bench_8gpu_chunk32.json 2> $OUT/${TAG}_bench_8gpu_chunk32.err; echo "chunk32 exit $?"
for f in $OUT/${TAG}_bench_8gpu.json $OUT/${TAG}_bench_8gpu_maxctas4.json $OUT/${TAG}_bench_8gpu_chunk32.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", d["e2e"]["value"], "gemm", round(d["roofline"]["achieved"],0), "ovl", d["roofline"].get("overlapped_with_allreduce",{}).get("vs_quiet_layers"), "dp", d["dp_check"]["status"], "comm", d["comm"]["comm_exposed_ms_per_step_by_rank"], "step_by_rank", d["comm"]["step_ms_by_rank"])
PY
done
