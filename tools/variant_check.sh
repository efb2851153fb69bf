#!/bin/bash
# GPU-box A/B of an FFT variant selected by an environment variable: parity first, then bench.
#   tools/variant_check.sh BGPU_FFT_RECUR [grid]
VAR=$1; GRID=${2:-256}
mkdir -p gpurun_out
env $VAR=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft_matches_numpy or tma_pass or large_grid or gradient_128" 2>&1 | tail -4
for v in 0 1; do
  env $VAR=$v timeout 300 python bench.py --grid $GRID --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/variant.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
pk = d['roofline']['per_kernel']
print('$VAR=$v grid $GRID: %.1f evals/s  %.3f ms  exact %.1f  e2e %.1f leap %.1f | ' % (d['value'], d['ms_per_step'], d['also']['gradient_evals_per_s_calc_h_4'], d['e2e']['value'], d['also']['leapfrog_steps_per_s']) + ' '.join('%s=%.3f' % (k, v['ms_per_step']) for k, v in pk.items()))
"
done
