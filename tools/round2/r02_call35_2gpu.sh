#!/bin/bash
# round 2, call 35 (2 GPUs): shared x pass on slabs (inverse direction): slab tests across two devices, 512^3 slab rate with / without
OUT=gpurun_out/r02_c35
mkdir -p $OUT
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_slab.log
for v in 1 0; do
  BGPU_SHARE_X_SLAB=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$v bench.py --gpus 2 --mode slab --grid 512 --steps 10 --warmup 3 > $OUT/slab512_sx$v.json 2> $OUT/slab512_sx$v.err
  echo "slab bench sx=$v rc=$?"
done
python - <<'PY'
import json
for v in ("1", "0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02_c35/slab512_sx{v}.json") if l.startswith("{")][-1])
        g = d.get("slab_detail") or d
        pk = d.get("per_kernel") or d.get("roofline", {}).get("per_kernel") or {}
        print("share_x_slab", v, "%.2f evals/s" % d["value"], d.get("leapfrog"), {k: round(x["ms_per_step"], 3) for k, x in pk.items()})
    except Exception as e:
        print("failed:", e)
PY
