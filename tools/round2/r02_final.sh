#!/bin/bash
# round 2, final check: what the driver runs at round end -- the GPU suite, smoke(), the default bench, the reference arm
OUT=gpurun_out/r02_final
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 2>&1 | tail -14 | tee $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
( time timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2> $OUT/bench_default.time
echo "bench default rc=$?"; tail -3 $OUT/bench_default.time
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err ) 2> $OUT/bench_reference.time
echo "bench reference rc=$?"; tail -3 $OUT/bench_reference.time
timeout 300 python bench.py --mode slab --grid 256 --steps 5 --warmup 2 > $OUT/slab256_1gpu.json 2> $OUT/slab256_1gpu.err
echo "slab mode on one GPU rc=$?"; python -c "import json; d=json.loads(open('$OUT/slab256_1gpu.json').readline()); print('slab 256 on 1 GPU: %.1f evals/s' % d['value'], d.get('leapfrog'))"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final/bench_default.json").readline())
r = json.loads(open("gpurun_out/r02_final/bench_reference.json").readline())
e = d["e2e"]
print("value %.1f  e2e %.1f  reference %.3f  e2e ratio %.1f" % (d["value"], e["value"], r["value"], e["value"] / r["value"]))
print("whole", d["roofline"]["whole_path"]["frac"], d["roofline"]["whole_path_exact_adjoint"]["frac"], "dominant", d["roofline"]["kernel"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("leapfrog", d["also"]["leapfrog_steps_per_s"], "candidate ms", d["also"]["hmc_candidate_ms_neps8_device_resident"], "e2e traj", e["trajectory"]["leapfrog_steps_per_s"], "cand", e["candidate"]["leapfrog_steps_per_s"], "chains", e["interleaved_chains"].get("value"))
g = d["also"]["grid_512"]
print("512:", g.get("gradient_evals_per_s"), g.get("exact_adjoint"), g.get("leapfrog_steps_per_s"), g.get("e2e_evals_per_s"), g.get("error"))
print("sph", d["also"]["sph_default_config"]["gradient_evals_per_s"], "f32", d["also"]["single_precision_mode"]["gradient_evals_per_s"], d["also"]["single_precision_mode"]["leapfrog_steps_per_s"])
print("clocks", d["clocks"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"])
PY
