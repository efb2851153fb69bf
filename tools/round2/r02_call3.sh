#!/bin/bash
# round 2, call 3: ncu of the particle kernels (sweep vs first generation), fused leapfrog through the GPU suite, bench
OUT=gpurun_out/r02_c3
mkdir -p $OUT
timeout 300 tools/native/grad_ab BGPU_SWEEP=0 256 4 > $OUT/plain_grad_ab.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scatter|gather" -c 10 -o $OUT/prof_particles -f \
   tools/native/grad_ab BGPU_SWEEP=0 256 4 > $OUT/ncu_particles.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_particles.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
BGPU_LEAPFROG_FUSED=0 timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256_unfused.json 2> $OUT/bench256_unfused.err
python - <<'PY'
import json
for tag in ("", "_unfused"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c3/bench256{tag}.json").readline())
        print(tag or "fused", "%.1f evals/s" % d["value"], {k: v for k, v in d["also"].items() if "leapfrog" in k or "calc_h_4" in k}, "e2e", d["e2e"].get("value"), d["e2e"].get("trajectory"))
    except Exception as e:
        print(tag, "failed:", e)
PY
