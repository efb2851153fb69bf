#!/bin/bash
# First GPU call of round 2: run the experiments that were written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Everything lands in gpurun_out/r02_first/.  Needs the prebuilt libbarcode_b200.so and tools/native/{fft_ab,grad_ab}
# (g++ -O2 -fopenmp -I include tools/native/X.cc -L barcode_b200 -lbarcode_b200 -Wl,-rpath,'$ORIGIN/../../barcode_b200' -o tools/native/X).
OUT=gpurun_out/r02_first
mkdir -p $OUT
for t in fft_ab grad_ab; do   # the binaries are git-ignored: rebuild them where they are missing (g++ is on the box)
  [ -x tools/native/$t ] || g++ -O2 -fopenmp -I include tools/native/$t.cc -L barcode_b200 -lbarcode_b200 \
      -Wl,-rpath,'$ORIGIN/../../barcode_b200' -o tools/native/$t
done
# 1. shared x pass: parity + per-kernel times, Python-free, seconds each
for cfg in "256 0" "256 4" "128 0 3 0" "128 4 3 0" "128 0 1 0 0" "512 0"; do
  timeout 120 tools/native/grad_ab BGPU_SHARE_X $cfg > "$OUT/grad_ab_share_x_${cfg// /_}.log" 2>&1
  tail -4 "$OUT/grad_ab_share_x_${cfg// /_}.log"
done
# 1b. memcheck of the extended-functor kernels at a small grid
for v in BGPU_SHARE_X; do   # (BGPU_FFT_2WARP only changes N = 512: too slow under memcheck)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 tools/native/grad_ab $v 128 0 > "$OUT/memcheck_$v.log" 2>&1
  echo "memcheck $v: exit $?"; tail -2 "$OUT/memcheck_$v.log"
done
# 2. its gated parity tests
BGPU_UNVERIFIED_TESTS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "shared_x" 2>&1 | tail -5 | tee $OUT/pytest_shared_x.log
# 2b. the reference's own integration smoke configuration (8^3, SPH, calc_h = 2) through the drop-in
BGPU_UNVERIFIED_TESTS=1 timeout 600 python -m pytest tests/test_dropin.py -m gpu -q -k "smoke_config" 2>&1 | tail -5 | tee $OUT/pytest_smoke_config.log
# 3. bench with and without it (the default line also carries e2e.interleaved_chains for the first time)
timeout 600 python bench.py --grid 256 --no-cpu-baseline > $OUT/bench256_default.json 2> $OUT/bench256_default.err
BGPU_SHARE_X=1 timeout 600 python bench.py --grid 256 --no-cpu-baseline --no-e2e-chains > $OUT/bench256_share_x.json 2> $OUT/bench256_share_x.err
python - <<'PY'
import json
for tag in ("default", "share_x"):
    try:
        d = json.loads(open(f"gpurun_out/r02_first/bench256_{tag}.json").readline())
        pk = d["roofline"]["per_kernel"]
        print(tag, "%.1f evals/s" % d["value"], "exact %.1f" % d["also"]["gradient_evals_per_s_calc_h_4"],
              "e2e %.1f" % d["e2e"]["value"], "interleaved", d["e2e"].get("interleaved_chains"),
              " ".join("%s=%.3f/%g" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in pk.items()))
    except Exception as e:
        print(tag, "failed:", e)
PY
