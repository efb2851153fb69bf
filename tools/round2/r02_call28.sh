#!/bin/bash
# round 2, call 28: leapfrog-form tests; ncu launch list of the final bench command; ncu --set full of one 512^3 evaluation
OUT=gpurun_out/r02_c28
mkdir -p $OUT /tmp/ncu
timeout 600 python -m pytest tests/test_leapfrog_forms_gpu.py -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_forms.log
CMD="python bench.py --grid 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-chains --no-sph --no-f32 --no-512"
timeout 600 $CMD > $OUT/plain.log 2> $OUT/plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_bench256.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 tools/native/grad_ab BGPU_NOVARIANT 512 0 > $OUT/plain512.log 2>&1; echo "plain 512 rc=$?"; tail -4 $OUT/plain512.log | cut -c1-400
timeout 1200 ncu --set full --clock-control none -k regex:"fft_|scatter_cic|overdens" -s 83 -c 32 -o /tmp/ncu/prof512 -f \
   tools/native/grad_ab BGPU_NOVARIANT 512 0 > $OUT/ncu_512.log 2>&1
echo "ncu 512 rc=$?"
python tools/ncu_summary.py full /tmp/ncu/prof512.ncu-rep --json $OUT/traffic_r02_512.json > $OUT/ncu_full_r02_eval512.txt 2>&1
grep -c "launches; first shown" $OUT/ncu_full_r02_eval512.txt
grep "launches; first shown\|gpu__time_duration.sum\|dram read" $OUT/ncu_full_r02_eval512.txt | cut -c1-110
du -sh $OUT
