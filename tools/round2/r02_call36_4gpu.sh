#!/bin/bash
# round 2, call 36 (4 GPUs): slab chain with the shared x pass against the single-GPU chain at 128^3 / 256^3, then the 512^3 rate
OUT=gpurun_out/r02_c36
mkdir -p $OUT
for N in 128 256; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2961$((N/128)) tools/slab_check.py --grid $N > $OUT/slab_check_$N.log 2>&1
  echo "slab_check $N rc=$?"; tail -4 $OUT/slab_check_$N.log | cut -c1-300
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 4 --mode slab --grid 512 --steps 10 --warmup 3 > $OUT/slab512_4gpu.json 2> $OUT/slab512_4gpu.err
echo "slab bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_c36/slab512_4gpu.json") if l.startswith("{")][-1])
    print("4 GPUs 512^3: %.2f evals/s" % d["value"], d.get("leapfrog"), d.get("nvlink"), {k: round(x["ms_per_step"], 3) for k, x in d["per_kernel"].items()})
except Exception as e:
    print("failed:", e)
PY
