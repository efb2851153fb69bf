#!/bin/bash
# round 2, call 15: ncu of the particle kernels and of one evaluation's FFT passes, summarised ON the box
# (the .ncu-rep files are too big to travel back: 2 MB per kernel with sources)
OUT=gpurun_out/r02_c15
mkdir -p $OUT /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_cic|gather_cic|overdens" -c 3 -o /tmp/ncu/prof_particles -f \
   tools/native/grad_ab BGPU_NOVARIANT 256 4 > $OUT/ncu_particles.log 2>&1
echo "ncu particles rc=$?"
python tools/ncu_summary.py full /tmp/ncu/prof_particles.ncu-rep > $OUT/ncu_full_r02_particles_256.txt 2>&1
ncu -i /tmp/ncu/prof_particles.ncu-rep --page raw --csv > $OUT/ncu_raw_r02_particles_256.csv 2>/dev/null
ncu -i /tmp/ncu/prof_particles.ncu-rep --page source --csv -k regex:scatter_cic > $OUT/ncu_source_r02_scatter_256.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"fft_" -s 60 -c 36 -o /tmp/ncu/prof_fft256 -f \
   tools/native/grad_ab BGPU_NOVARIANT 256 0 > $OUT/ncu_fft.log 2>&1
echo "ncu fft rc=$?"
python tools/ncu_summary.py full /tmp/ncu/prof_fft256.ncu-rep --json $OUT/traffic_r02_fft256.json > $OUT/ncu_full_r02_fft256.txt 2>&1
CMD="python bench.py --grid 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-chains --no-sph"
timeout 600 $CMD > $OUT/plain.log 2> $OUT/plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench256.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
du -sh $OUT
