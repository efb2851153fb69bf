#!/bin/bash
# round 2, call 7 (2 GPUs): slab tests (fused leapfrog + run-away guard on slabs), bench --gpus 2 with the slab leg
OUT=gpurun_out/r02_c7
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_slab.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
echo "bench rc=$?"; tail -3 $OUT/bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_c7/bench_2gpu.json").readline())
    print("chains %.1f evals/s" % d["value"], "e2e %.1f" % d["e2e"]["value"], d["e2e"].get("numa"))
    s = d["slab"]
    print("slab parity", s.get("parity_max_rel"), "error", s.get("error"))
    for g in ("512", "1024"):
        if g in s and "value" in s[g]:
            print(g, "%.2f evals/s" % s[g]["value"], s[g]["nvlink"], {k: round(v["ms_per_step"], 3) for k, v in s[g]["per_kernel"].items()})
        elif g in s:
            print(g, s[g])
except Exception as e:
    print("failed:", e)
PY
