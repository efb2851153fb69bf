#!/bin/bash
# round 2, call 19: single-precision mode after the measurements of call 18 (CTA-barrier strided pass kept, table twiddles,
# warp-synchronous z passes): tests, bench line with 16- and 8-pencil tiles
OUT=gpurun_out/r02_c19
mkdir -p $OUT
timeout 900 python -m pytest tests/test_f32_gpu.py -m gpu -q -s 2>&1 | grep -E "f32 vs fp64|empty-cell|f32 GPU|reference SINGLE|passed|failed|Error|error" | cut -c1-330 | tee $OUT/pytest_f32.log
for t in 16 8; do
  BGPU_F32_T=$t timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains --no-sph > $OUT/bench256_T$t.json 2> $OUT/bench256_T$t.err
  echo "bench T=$t rc=$?"
done
BGPU_F32_T=8 timeout 300 python -m pytest tests/test_f32_gpu.py -m gpu -q -k "matches_the_fp64_path and (64 or 128)" 2>&1 | tail -2
python - <<'PY'
import json
for t in ("16", "8"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c19/bench256_T{t}.json").readline())
        print("T", t, "fp64 %.1f evals/s" % d["value"], json.dumps(d["also"]["single_precision_mode"]))
    except Exception as e:
        print("failed:", e)
PY
