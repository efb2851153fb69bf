#!/bin/bash
OUT=gpurun_out/r02_c10
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $OUT/pytest_gpu.log
timeout 900 python bench.py --grid 256 > $OUT/bench256.json 2> $OUT/bench256.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_c10/bench256.json").readline())
    print("%.1f evals/s" % d["value"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["also"].items() if k != "sph_default_config"})
    print("sph", d["also"]["sph_default_config"])
    print("e2e", {k: v for k, v in d["e2e"].items() if k in ("value", "candidate", "numa")})
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("failed:", e)
PY
