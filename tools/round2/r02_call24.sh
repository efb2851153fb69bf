#!/bin/bash
# round 2, call 24: f32 strided pass with two tiles in flight (three buffers) against one
OUT=gpurun_out/r02_c24
mkdir -p $OUT
for v in 2 1; do
  echo "BGPU_F32_PD=$v"
  BGPU_F32_PD=$v timeout 300 python tools/f32_times.py --grid 256 --steps 3 2>&1 | tail -9 | tee $OUT/f32_times_pd$v.log
done
BGPU_F32_PD=2 timeout 300 python tools/f32_times.py --grid 512 --steps 2 2>&1 | tail -9 | tee $OUT/f32_times512_pd2.log
BGPU_F32_PD=1 timeout 300 python tools/f32_times.py --grid 512 --steps 2 2>&1 | tail -9 | tee $OUT/f32_times512_pd1.log
timeout 900 python -m pytest tests/test_f32_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_f32.log
