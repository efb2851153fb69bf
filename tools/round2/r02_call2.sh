#!/bin/bash
# round 2, call 2: x-sweep scatter / gather A/B (variant = BGPU_SWEEP=0, the first-generation kernels), then the GPU suite
OUT=gpurun_out/r02_c2
mkdir -p $OUT
for cfg in "256 0" "256 4" "128 0 1 0 0 2" "128 4 1 0 1 2" "512 4" "64 4 1 0"; do
  timeout 180 tools/native/grad_ab BGPU_SWEEP=0 $cfg > "$OUT/grad_ab_sweep_${cfg// /_}.log" 2>&1
  tail -5 "$OUT/grad_ab_sweep_${cfg// /_}.log"
done
for seg in 16 64 256; do
  echo "seg $seg"; timeout 120 tools/native/grad_ab BGPU_SWEEP=$seg 256 4 2>&1 | grep -E "variant|relative"
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256.json 2> $OUT/bench256.err; tail -c 1500 $OUT/bench256.json
