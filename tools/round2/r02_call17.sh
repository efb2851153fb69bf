#!/bin/bash
# round 2, call 17: first GPU run of the single-precision mode: its tests (with error printouts), memcheck at 32^3, bench line
OUT=gpurun_out/r02_c17
mkdir -p $OUT
timeout 900 python -m pytest tests/test_f32_gpu.py -m gpu -q -s -x 2>&1 | grep -vE "^\s*$" | tail -60 | tee $OUT/pytest_f32.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_f32_gpu.py -m gpu -q -x -k "32-1-True-0 or 32-0-False-4" > $OUT/memcheck_f32.log 2>&1
echo "memcheck rc=$?"; tail -5 $OUT/memcheck_f32.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains --no-sph > $OUT/bench256.json 2> $OUT/bench256.err
echo "bench rc=$?"; tail -3 $OUT/bench256.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_c17/bench256.json").readline())
    print("fp64 %.1f evals/s" % d["value"], "calc_h4 %.1f" % d["also"]["gradient_evals_per_s_calc_h_4"])
    print(json.dumps(d["also"]["single_precision_mode"], indent=1))
except Exception as e:
    print("failed:", e)
PY
