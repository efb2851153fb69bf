#!/bin/bash
# round 2, call 25: the driver's bench command (with the 512^3 line), ncu of the final SPH kernels
OUT=gpurun_out/r02_c25
mkdir -p $OUT /tmp/ncu
( time timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2> $OUT/bench_default.time
echo "bench default rc=$?"; tail -3 $OUT/bench_default.time
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sph_cols5" -c 2 -o /tmp/ncu/prof_sph -f \
   tools/native/grad_ab BGPU_NOVARIANT 256 2 1 0 1 3 > $OUT/ncu_sph.log 2>&1
echo "ncu sph rc=$?"
python tools/ncu_summary.py full /tmp/ncu/prof_sph.ncu-rep > $OUT/ncu_full_r02_sph_z5_256.txt 2>&1
cat $OUT/ncu_full_r02_sph_z5_256.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c25/bench_default.json").readline())
print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), "whole", d["roofline"]["whole_path"]["frac"], d["roofline"].get("whole_path_exact_adjoint"))
print("512:", json.dumps(d["also"]["grid_512"])[:1500])
print("sph:", json.dumps(d["also"]["sph_default_config"])[:600])
print("cpu:", d["cpu_baseline"])
PY
