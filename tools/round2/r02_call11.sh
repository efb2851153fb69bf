#!/bin/bash
OUT=gpurun_out/r02_c11
mkdir -p $OUT
timeout 900 python -m pytest tests/test_slab_local_gpu.py -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_local_slab.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_slab_local_gpu.py 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
timeout 900 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_c11/bench256.json").readline())
    print("%.1f evals/s" % d["value"], "sph", d["also"]["sph_default_config"])
except Exception as e:
    print("failed:", e)
PY
