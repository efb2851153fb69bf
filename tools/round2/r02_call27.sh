#!/bin/bash
# round 2, call 27: leapfrog in k-space: the whole GPU suite, leapfrog rate against the fused real-space form
OUT=gpurun_out/r02_c27
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -16 | tee $OUT/pytest_gpu.log
for v in 1 0; do
  BGPU_LEAPFROG_KSPACE=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e-chains --no-512 --no-f32 --no-sph > $OUT/bench256_k$v.json 2> $OUT/bench256_k$v.err
  echo "bench kspace=$v rc=$?"
done
BGPU_LEAPFROG_KSPACE=1 timeout 600 python bench.py --grid 512 --steps 5 --no-cpu-baseline --no-e2e-chains --no-512 --no-f32 --no-sph > $OUT/bench512_k1.json 2> $OUT/bench512_k1.err
python - <<'PY'
import json
for t in ("256_k1", "256_k0", "512_k1"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c27/bench{t}.json").readline())
        e = d["e2e"]
        print(t, "grad %.1f" % d["value"], "leapfrog %.1f" % d["also"]["leapfrog_steps_per_s"], "cand ms %.2f" % d["also"]["hmc_candidate_ms_neps8_device_resident"],
              "traj e2e steps/s %.1f" % e["trajectory"]["leapfrog_steps_per_s"], "cand steps/s %.1f" % e["candidate"]["leapfrog_steps_per_s"])
    except Exception as ex:
        print(t, "failed:", ex)
PY
