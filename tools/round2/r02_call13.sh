#!/bin/bash
# round 2, call 13: lean particle kernels and ping-pong traversal A/B, SPH + host-stream momenta on slabs, GPU suite, bench
OUT=gpurun_out/r02_c13
mkdir -p $OUT
for cfg in "256 0" "256 4" "512 4"; do
  timeout 180 tools/native/grad_ab BGPU_LEAN=0 $cfg > "$OUT/grad_ab_lean_${cfg// /_}.log" 2>&1
  grep -E "gradient_psi:|relative|FAILED" "$OUT/grad_ab_lean_${cfg// /_}.log" | sed -E 's/fft_[a-z_0-9]+ [0-9.]+ ms \/ [0-9]+ //g'
done
for cfg in "256 0" "256 4"; do
  timeout 180 tools/native/grad_ab BGPU_PINGPONG=1 $cfg > "$OUT/grad_ab_pingpong_${cfg// /_}.log" 2>&1
  grep -E "gradient_psi:|relative|FAILED" "$OUT/grad_ab_pingpong_${cfg// /_}.log"
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
BGPU_PINGPONG=1 timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains --no-sph > $OUT/bench256_pingpong.json 2> $OUT/bench256_pingpong.err
python - <<'PY'
import json
for tag in ("256", "256_pingpong"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c13/bench{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["also"].items() if not isinstance(v, dict)},
              "whole %.3f" % d["roofline"]["whole_path"]["frac"],
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
