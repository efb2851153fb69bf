#!/bin/bash
# round 2, call 5: z round-trip kernel A/B (variant = BGPU_ZROUND=0), parity, bench
OUT=gpurun_out/r02_c5
mkdir -p $OUT
for cfg in "256 0" "512 0" "128 0"; do
  timeout 180 tools/native/grad_ab BGPU_ZROUND=0 $cfg > "$OUT/grad_ab_zround_${cfg// /_}.log" 2>&1
  grep -E "gradient_psi:|relative|FAILED" "$OUT/grad_ab_zround_${cfg// /_}.log"
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
python - <<'PY'
import json
for tag in ("256",):
    try:
        d = json.loads(open(f"gpurun_out/r02_c5/bench{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], {k: round(v, 1) for k, v in d["also"].items() if "leapfrog" in k or "calc_h_4" in k},
              "e2e %.1f" % d["e2e"]["value"], "whole %.3f" % d["roofline"]["whole_path"]["frac"],
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
