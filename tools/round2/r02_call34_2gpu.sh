#!/bin/bash
# round 2, call 34 (2 GPUs): slab tests across two devices with the k-space leapfrog, N = 2 bench (chains + slab leg with its leapfrog rate)
OUT=gpurun_out/r02_c34
mkdir -p $OUT
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_slab.log
for v in 1 0; do
  BGPU_LEAPFROG_KSPACE=$v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --steps 10 --warmup 3 --no-sph --no-f32 > $OUT/bench_2gpu_k$v.json 2> $OUT/bench_2gpu_k$v.err
  echo "bench k=$v rc=$?"
done
python - <<'PY'
import json
for v in ("1", "0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02_c34/bench_2gpu_k{v}.json") if l.startswith("{")][-1])
        s = d["slab"]
        print("kspace", v, "chains %.1f" % d["value"], "slab parity", s.get("parity_max_rel"), "err", s.get("error"))
        g = s["512"]
        print("   512: %.2f evals/s" % g["value"], g.get("leapfrog"), g.get("nvlink", {}).get("GBps_sent_per_gpu"))
    except Exception as e:
        print("failed:", e)
PY
