#!/bin/bash
# round 2, call 21: ncu --set full of the two SPH kernels (static column lists) at 256^3
OUT=gpurun_out/r02_c21
mkdir -p $OUT /tmp/ncu
timeout 300 tools/native/grad_ab BGPU_NOVARIANT 256 2 1 0 1 3 > $OUT/plain.log 2>&1; echo "plain rc=$?"; tail -5 $OUT/plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sph_cols" -c 2 -o /tmp/ncu/prof_sph -f \
   tools/native/grad_ab BGPU_NOVARIANT 256 2 1 0 1 3 > $OUT/ncu_sph.log 2>&1
echo "ncu sph rc=$?"
python tools/ncu_summary.py full /tmp/ncu/prof_sph.ncu-rep > $OUT/ncu_full_r02_sph_256.txt 2>&1
ncu -i /tmp/ncu/prof_sph.ncu-rep --page raw --csv > $OUT/ncu_raw_r02_sph_256.csv 2>/dev/null
ncu -i /tmp/ncu/prof_sph.ncu-rep --page source --csv -k regex:scatter_sph > $OUT/ncu_source_r02_sph_scatter.csv 2>/dev/null
ncu -i /tmp/ncu/prof_sph.ncu-rep --page source --csv -k regex:gather_sph > $OUT/ncu_source_r02_sph_gather.csv 2>/dev/null
cat $OUT/ncu_full_r02_sph_256.txt
du -sh $OUT
