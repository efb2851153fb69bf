#!/bin/bash
# round 2, call 16: scatter / gather occupancy + read-ahead variants A/B, the whole GPU suite, the driver's bench command, 512^3 bench
OUT=gpurun_out/r02_c16
mkdir -p $OUT
for v in 22 32 33 43; do
  echo "BGPU_LEAN=$v"
  timeout 180 tools/native/grad_ab BGPU_LEAN=$v 256 4 > "$OUT/grad_ab_lean$v.log" 2>&1
  grep -E "gradient_psi:|relative|FAILED" "$OUT/grad_ab_lean$v.log" | grep -oE "(base|variant)[^:]*:|(scatter|gather_adjoint) [0-9.]+ ms|relative.*|FAILED.*"
done
timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -30 | tee $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench default rc=$?"
timeout 600 python bench.py --grid 512 --no-cpu-baseline --no-e2e-chains --no-sph > $OUT/bench512.json 2> $OUT/bench512.err
echo "bench 512 rc=$?"
python - <<'PY'
import json
for tag in ("_default", "512"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c16/bench{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], "e2e", d["e2e"]["value"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["also"].items() if not isinstance(v, dict)},
              "whole %.3f" % d["roofline"]["whole_path"]["frac"],
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
