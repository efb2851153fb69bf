tools/slab_scaling.sh 8 512 1024
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/chains_256_8gpu.json 2> gpurun_out/chains_256_8gpu.err
echo "chains rc=$?"; head -c 400 gpurun_out/chains_256_8gpu.json
