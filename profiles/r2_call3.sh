#!/bin/bash
# round-2 GPU call 3 (2 GPUs): NCCL paths behind the C-ABI + C5 strong-sharded over 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_c1_shape.py tests/test_synth.py -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2_multigpu_tests.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
   > gpurun_out/r2_bench_c5_2gpu_a.json 2> gpurun_out/r2_bench_c5_2gpu_a.err; echo "c5x2 rc=$?"; tail -c 2500 gpurun_out/r2_bench_c5_2gpu_a.json; tail -5 gpurun_out/r2_bench_c5_2gpu_a.err
