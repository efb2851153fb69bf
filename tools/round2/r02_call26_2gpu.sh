#!/bin/bash
# round 2, call 26 (2 GPUs): slab GPU tests across two devices, the driver's N = 2 bench command (chains + slab leg)
OUT=gpurun_out/r02_c26
mkdir -p $OUT
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_slab.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err ) 2> $OUT/bench_2gpu.time
echo "bench rc=$?"; tail -3 $OUT/bench_2gpu.err; tail -3 $OUT/bench_2gpu.time
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_c26/bench_2gpu.json") if l.startswith("{")][-1])
    print("chains %.1f evals/s" % d["value"], "e2e %.1f" % d["e2e"]["value"], d["e2e"].get("numa"))
    s = d["slab"]
    print("slab parity", s.get("parity_max_rel"), "error", s.get("error"))
    for g in ("512", "1024"):
        if g in s and isinstance(s[g], dict) and "value" in s[g]:
            print(g, "%.2f evals/s" % s[g]["value"], s[g].get("nvlink"), {k: round(v["ms_per_step"], 3) for k, v in s[g]["per_kernel"].items()})
        elif g in s:
            print(g, s[g])
except Exception as e:
    print("failed:", e)
PY
