#!/bin/bash
# round-2 GPU call 20 (2 GPUs): NCCL tests on the final tree, C5 strong-sharded over 2 GPUs (re-aligned pairs + phased remainder per shard)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multigpu_tests_final.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/r2_multigpu_tests_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/r2_bench_c5_2gpu_phased.json 2> gpurun_out/r2_bench_c5_2gpu_phased.err; echo "c5 2gpu rc=$?"; tail -c 2600 gpurun_out/r2_bench_c5_2gpu_phased.json; tail -3 gpurun_out/r2_bench_c5_2gpu_phased.err
