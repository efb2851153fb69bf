#!/bin/bash
# round 2, call 6: PDL A/B (variant = BGPU_PDL=0), the 512^3 reference case, bench 256 / 512
OUT=gpurun_out/r02_c6
mkdir -p $OUT
for cfg in "256 0" "256 4" "512 0"; do
  timeout 180 tools/native/grad_ab BGPU_PDL=0 $cfg > "$OUT/grad_ab_pdl_${cfg// /_}.log" 2>&1
  grep -E "relative|FAILED" "$OUT/grad_ab_pdl_${cfg// /_}.log"
done
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
BGPU_PDL=0 timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256_nopdl.json 2> $OUT/bench256_nopdl.err
timeout 600 python bench.py --grid 512 --steps 5 --no-cpu-baseline --no-e2e-chains > $OUT/bench512.json 2> $OUT/bench512.err
BGPU_PDL=0 timeout 600 python bench.py --grid 512 --steps 5 --no-cpu-baseline --no-e2e-chains > $OUT/bench512_nopdl.json 2> $OUT/bench512_nopdl.err
python - <<'PY'
import json
for tag in ("256", "256_nopdl", "512", "512_nopdl"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c6/bench{tag}.json").readline())
        print(tag, "%.1f evals/s" % d["value"], {k: round(v, 1) for k, v in d["also"].items() if "leapfrog" in k or "calc_h_4" in k or "candidate" in k},
              "e2e %.1f" % d["e2e"]["value"], "whole %.3f" % d["roofline"]["whole_path"]["frac"])
    except Exception as e:
        print(tag, "failed:", e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
