#!/bin/bash
# One point of the strong-scaling curve of the slab-decomposed chain (run under gpurun --gpus G):
#   tools/slab_scaling.sh G [grids...]
G=$1; shift
GRIDS=${@:-"512 1024"}
mkdir -p gpurun_out
for N in $GRIDS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus $G --mode slab --grid $N --steps 3 --warmup 2 > gpurun_out/slab_${N}_${G}gpu.json 2> gpurun_out/slab_${N}_${G}gpu.err
  echo "slab $N on $G GPUs rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/slab_${N}_${G}gpu.json").readline())
    print("  %.2f evals/s  %.1f ms/step" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 2) for k, v in d["per_kernel"].items()})
except Exception as e:
    print("  no result:", e)
PY
done
