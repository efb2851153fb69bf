#!/bin/bash
# GPU-box check of the fused z+y kernel: parity first (under a timeout: the roles wait on each other), then speed
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft_matches_numpy or tma_pass or large_grid" 2>&1 | tail -15
echo "pytest rc=$?"
for fu in 0 1; do
  for lead in 24 48 96; do
    if [ $fu = 0 ] && [ $lead != 48 ]; then continue; fi
    BGPU_FFT_FUSED=$fu BGPU_FFT_LEAD=$lead timeout 300 python bench.py --grid 256 --steps 10 --warmup 3 --no-cpu-baseline \
      2>gpurun_out/fused.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
pk = d['roofline']['per_kernel']
print('fused $fu lead $lead : %.1f evals/s  %.3f ms  exact %.1f  e2e %.1f leap %.1f | ' % (d['value'], d['ms_per_step'], d['also']['gradient_evals_per_s_calc_h_4'], d['e2e']['value'], d['also']['leapfrog_steps_per_s']) + ' '.join('%s=%.3f' % (k, v['ms_per_step']) for k, v in pk.items()))
"
  done
done
