#!/bin/bash
# round 2, call 4: sweep A/B on a smooth field, ncu of the particle kernels on it, GPU suite with the exact run-away guard, bench
OUT=gpurun_out/r02_c4
mkdir -p $OUT
for cfg in "256 0" "256 4" "512 4" "128 4"; do
  timeout 180 tools/native/grad_ab BGPU_SWEEP=0 $cfg > "$OUT/grad_ab_sweep_${cfg// /_}.log" 2>&1
  grep -E "gradient_psi:|relative" "$OUT/grad_ab_sweep_${cfg// /_}.log" | sed -E 's/fft_[a-z_0-9]+ [0-9.]+ ms \/ [0-9]+ //g'
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scatter|gather" -c 10 -o $OUT/prof_particles -f \
   tools/native/grad_ab BGPU_SWEEP=0 256 4 > $OUT/ncu_particles.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --grid 256 --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
timeout 600 python bench.py --grid 512 --steps 5 --no-cpu-baseline --no-e2e-chains > $OUT/bench512.json 2> $OUT/bench512.err
python - <<'PY'
import json
for tag in ("256", "512"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c4/bench{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], {k: round(v, 1) for k, v in d["also"].items() if "leapfrog" in k or "calc_h_4" in k},
              "e2e %.1f" % d["e2e"]["value"], "whole %.3f" % d["roofline"]["whole_path"]["frac"],
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
