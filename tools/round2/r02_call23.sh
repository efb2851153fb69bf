#!/bin/bash
# round 2, call 23: SPH kernels with the unrolled z loop (K <= 2): parity tests, timing against the list kernels
OUT=gpurun_out/r02_c23
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "sph or dropin or assign or smoke" 2>&1 | tail -8 | tee $OUT/pytest_sph.log
timeout 300 tools/native/grad_ab BGPU_SPH_Z5=0 256 2 1 0 1 3 2>&1 | tail -5 | tee $OUT/ab256.log
timeout 300 tools/native/grad_ab BGPU_SPH_Z5=0 128 2 1 1 1 3 2>&1 | tail -4 | tee $OUT/ab128rsd.log
