#!/bin/bash
# round 2, call 18: single-precision mode with the warp-synchronous passes: tests (both strided-pass forms), bench line
OUT=gpurun_out/r02_c18
mkdir -p $OUT
timeout 900 python -m pytest tests/test_f32_gpu.py -m gpu -q -s 2>&1 | grep -E "f32 vs fp64|f32 GPU|reference SINGLE|passed|failed|Error|error" | cut -c1-330 | tee $OUT/pytest_f32.log
BGPU_F32_WARP=0 timeout 900 python -m pytest tests/test_f32_gpu.py -m gpu -q -k "matches_the_fp64_path" 2>&1 | tail -3 | tee $OUT/pytest_f32_nowarp.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains --no-sph > $OUT/bench256.json 2> $OUT/bench256.err
echo "bench rc=$?"; tail -3 $OUT/bench256.err
timeout 600 python bench.py --grid 128 --no-cpu-baseline --no-e2e-chains --no-sph > $OUT/bench128.json 2> $OUT/bench128.err
python - <<'PY'
import json
for g in ("256", "128"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c18/bench{g}.json").readline())
        print(g, "fp64 %.1f evals/s" % d["value"], "calc_h4 %.1f" % d["also"]["gradient_evals_per_s_calc_h_4"])
        print(json.dumps(d["also"]["single_precision_mode"]))
    except Exception as e:
        print("failed:", e)
PY
