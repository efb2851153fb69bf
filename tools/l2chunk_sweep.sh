#!/bin/bash
# sweep of the L2-resident z+y sweep's chunk size / stream count (fft_plan.cu); run on the GPU box
mkdir -p gpurun_out
for g in 256 512; do
  for mb in 0 16 32 64; do
    for ns in 1 2; do
      if [ $mb = 0 ] && [ $ns = 1 ]; then continue; fi
      BGPU_FFT_L2CHUNK_MB=$mb BGPU_FFT_L2STREAMS=$ns python bench.py --grid $g --steps 10 --warmup 3 --no-cpu-baseline \
        2>gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
pk = d['roofline']['per_kernel']
print('grid $g mb $mb streams $ns : %.1f evals/s  %.3f ms  exact %.1f  e2e %.1f | ' % (d['value'], d['ms_per_step'], d['also']['gradient_evals_per_s_calc_h_4'], d['e2e']['value']) + ' '.join('%s=%.3f' % (k, v['ms_per_step']) for k, v in pk.items()))
"
    done
  done
done
