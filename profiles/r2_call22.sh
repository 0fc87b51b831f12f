#!/bin/bash
# round-2 GPU call 22 (8 GPUs): C5 strong-sharded over 8 GPUs on the final tree (re-aligned pairs, phased remainder, sqrt seed sample per shard)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 \
  > gpurun_out/r2_bench_c5_8gpu_phased.json 2> gpurun_out/r2_bench_c5_8gpu_phased.err; echo "c5 8gpu rc=$?"; tail -c 2600 gpurun_out/r2_bench_c5_8gpu_phased.json; tail -3 gpurun_out/r2_bench_c5_8gpu_phased.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 6 --warmup 3 --skip-parity --skip-cpu \
  > gpurun_out/r2_bench_c5_4gpu_phased.json 2> gpurun_out/r2_bench_c5_4gpu_phased.err; echo "c5 4gpu rc=$?"; tail -c 900 gpurun_out/r2_bench_c5_4gpu_phased.json
