#!/bin/bash
OUT=gpurun_out/r02_c8
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $OUT/pytest_gpu.log
