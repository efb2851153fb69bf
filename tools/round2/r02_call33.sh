#!/bin/bash
# round 2, call 33: f32 leapfrog in k-space: tests and rates
OUT=gpurun_out/r02_c33
mkdir -p $OUT
timeout 900 python -m pytest tests/test_f32_gpu.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_f32.log
for v in 1 0; do
  BGPU_LEAPFROG_KSPACE=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e-chains --no-512 --no-sph > $OUT/bench256_k$v.json 2> $OUT/bench256_k$v.err
  echo "bench kspace=$v rc=$?"
done
python - <<'PY'
import json
for t in ("1", "0"):
    d = json.loads(open(f"gpurun_out/r02_c33/bench256_k{t}.json").readline())
    f = d["also"]["single_precision_mode"]
    print("kspace", t, "fp64 grad %.1f leapfrog %.1f" % (d["value"], d["also"]["leapfrog_steps_per_s"]), "f32 grad %.1f leapfrog %.1f" % (f["gradient_evals_per_s"], f["leapfrog_steps_per_s"]), "clocks", d["clocks"])
PY
