#!/bin/bash
# round 2, call 29: candidate with cached psi of the current signal: tests (candidate, drop-in, leapfrog forms), candidate time
OUT=gpurun_out/r02_c29
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "candidate or dropin or forms or leapfrog" 2>&1 | tail -6 | tee $OUT/pytest.log
for v in 1 0; do
  BGPU_CANDIDATE_CACHE=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e-chains --no-512 --no-f32 --no-sph > $OUT/bench256_cache$v.json 2> $OUT/bench256_cache$v.err
  echo "bench cache=$v rc=$?"
done
python - <<'PY'
import json
for t in ("1", "0"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c29/bench256_cache{t}.json").readline())
        e = d["e2e"]
        print("cache", t, "grad %.1f" % d["value"], "leapfrog %.1f" % d["also"]["leapfrog_steps_per_s"], "cand ms %.2f" % d["also"]["hmc_candidate_ms_neps8_device_resident"],
              "cand steps/s %.1f" % e["candidate"]["leapfrog_steps_per_s"])
    except Exception as ex:
        print(t, "failed:", ex)
PY
