#!/bin/bash
# round 2, call 20: SPH scatter / gather over static column lists (particles_sph.cu): parity tests, A/B against the general kernels
OUT=gpurun_out/r02_c20
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "sph or dropin or assign or smoke" 2>&1 | tail -8 | tee $OUT/pytest_sph.log
for v in 1 0; do
  BGPU_SPH_COLS=$v timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains --no-f32 > $OUT/bench256_cols$v.json 2> $OUT/bench256_cols$v.err
  echo "bench cols=$v rc=$?"
done
python - <<'PY'
import json
for t in ("1", "0"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c20/bench256_cols{t}.json").readline())
        print("cols", t, "fp64 %.1f evals/s" % d["value"], json.dumps(d["also"]["sph_default_config"]))
    except Exception as e:
        print("failed:", e)
PY
