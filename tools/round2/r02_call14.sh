#!/bin/bash
# round 2, call 14: ncu of the particle kernels (lean sweep), segment-length sweep, launch list of the bench command
OUT=gpurun_out/r02_c14
mkdir -p $OUT
for seg in 8 32 64; do
  echo "seg $seg"; timeout 120 tools/native/grad_ab BGPU_SWEEP=$seg 256 4 2>&1 | grep -E "variant" | grep -oE "(scatter|gather_adjoint) [0-9.]+ ms"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_cic|gather_cic|overdens" -c 6 -o $OUT/prof_particles -f \
   tools/native/grad_ab BGPU_NOVARIANT 256 4 > $OUT/ncu_particles.log 2>&1
echo "ncu particles rc=$?"
CMD="python bench.py --grid 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-chains --no-sph"
timeout 600 $CMD > $OUT/plain.log 2> $OUT/plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench256.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
# one evaluation's FFT passes, full set: skip the set-up launches, 36 launches of the second evaluation
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fft_" -s 60 -c 36 -o $OUT/prof_fft256 -f \
   tools/native/grad_ab BGPU_NOVARIANT 256 0 > $OUT/ncu_fft.log 2>&1
echo "ncu fft rc=$?"
timeout 900 python -m pytest tests/test_slab_local_gpu.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_slab_local.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err
ls -la $OUT
python - <<'PY'
import json
for tag in ("256",):
    try:
        d = json.loads(open(f"gpurun_out/r02_c14/bench{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["also"].items() if not isinstance(v, dict)},
              "whole %.3f" % d["roofline"]["whole_path"]["frac"],
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
