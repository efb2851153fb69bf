#!/bin/bash
# GPU-box profile pass (one gpurun call): bench plain -> ncu launch list of the same command; one gradient
# evaluation plain -> ncu --set full of the FFT pass kernels.  Outputs under gpurun_out/, summarised here with
# tools/ncu_summary.py into profiles/.
TAG=${1:-r01d}
mkdir -p gpurun_out
CMD="python bench.py --grid 256 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>gpurun_out/plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "launch list rc=$?"
CMD2="python tools/kernel_times.py --grid 256 --steps 1"
$CMD2 > gpurun_out/plain_kt_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fft_strided_tma|fft_zpass_tma" -s 36 -c 36 -o gpurun_out/prof_fft256_$TAG -f $CMD2 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
python bench.py --grid 256 --steps 10 --warmup 3 > gpurun_out/bench256_$TAG.json 2>gpurun_out/bench256_$TAG.err
echo "bench256 rc=$?"
python bench.py --grid 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench512_$TAG.json 2>gpurun_out/bench512_$TAG.err
echo "bench512 rc=$?"
