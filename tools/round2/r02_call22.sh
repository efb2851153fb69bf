#!/bin/bash
# round 2, call 22: unrolled 5x5x5-box SPH kernels: parity tests, timing against the list kernels
OUT=gpurun_out/r02_c22
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "sph or dropin or assign or smoke" 2>&1 | tail -8 | tee $OUT/pytest_sph.log
for v in 1 0; do
  echo "BGPU_SPH_BOX=$v"
  BGPU_SPH_BOX=$v timeout 300 tools/native/grad_ab BGPU_NOVARIANT 256 2 1 0 1 3 2>&1 | grep -E "default gradient_psi|relative" | grep -oE "(scatter|gather_adjoint) [0-9.]+ ms|kernels total [0-9.]+ ms|relative.*"
done
BGPU_SPH_BOX=1 timeout 300 tools/native/grad_ab BGPU_SPH_BOX=0 256 2 1 0 1 3 2>&1 | tail -4
BGPU_SPH_BOX=1 timeout 300 tools/native/grad_ab BGPU_SPH_BOX=0 128 2 1 1 1 3 2>&1 | tail -4
# single-precision passes under ncu (which one is slow and why)
mkdir -p /tmp/ncu
timeout 300 python tools/f32_times.py --grid 256 --steps 2 2>&1 | tail -12 | tee $OUT/f32_times.log
timeout 600 ncu --set full --clock-control none -k regex:"zpass|strided_pass" -s 30 -c 12 -o /tmp/ncu/prof_f32 -f python tools/f32_times.py --grid 256 --steps 1 > $OUT/ncu_f32.log 2>&1
echo "ncu f32 rc=$?"
python tools/ncu_summary.py full /tmp/ncu/prof_f32.ncu-rep > $OUT/ncu_full_r02_f32_256.txt 2>&1
cat $OUT/ncu_full_r02_f32_256.txt | head -120
