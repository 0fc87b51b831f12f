#!/bin/bash
# round-2 GPU call 7 (8 GPUs): C5 strong-sharded over 8 and 4 GPUs + the multi-GPU tests with every device of the box
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 \
   > gpurun_out/r2_bench_c5_${n}gpu_a.json 2> gpurun_out/r2_bench_c5_${n}gpu_a.err; echo "c5x$n rc=$?"; tail -c 2600 gpurun_out/r2_bench_c5_${n}gpu_a.json; tail -3 gpurun_out/r2_bench_c5_${n}gpu_a.err
done
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs 2>&1 | tail -5 | tee gpurun_out/r2_multigpu_tests_8gpu.log
