#!/bin/bash
# round 2, call 12 (first call after the container was re-created): re-establish the round's evidence.
#   GPU suite, default bench (the driver's command), 512^3 bench, launch list of the bench command,
#   one ncu --set full capture of a whole evaluation at 256^3 (calc_h = 0 and 4) for per-kernel DRAM traffic.
OUT=gpurun_out/r02_c12
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench256.json 2> $OUT/bench256.err; echo "bench256 rc=$?"
timeout 900 python bench.py --grid 512 --steps 5 --no-cpu-baseline --no-e2e-chains > $OUT/bench512.json 2> $OUT/bench512.err; echo "bench512 rc=$?"
CMD="python bench.py --grid 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-chains --no-sph"
timeout 600 $CMD > $OUT/plain.log 2> $OUT/plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench256.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
for h in 0 4; do
  CMD2="tools/native/grad_ab BGPU_NOVARIANT 256 $h"
  timeout 120 $CMD2 > $OUT/grad_ab_256_$h.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -c 170 -o $OUT/prof_eval256_h$h -f $CMD2 > $OUT/ncu_full_h$h.log 2>&1
  echo "full capture h=$h rc=$?"
done
python - <<'PY'
import json
for tag in ("256", "512"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c12/bench{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["also"].items() if not isinstance(v, dict)},
              "e2e", {k: v for k, v in d["e2e"].items() if not isinstance(v, dict)}, "whole %.3f" % d["roofline"]["whole_path"]["frac"],
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
