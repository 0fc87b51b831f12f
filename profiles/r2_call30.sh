#!/bin/bash
# round-2 GPU call 30 (2 GPUs): NCCL tests + C5 over 2 GPUs on the final tree
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multigpu_tests_final.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r2_multigpu_tests_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 \
  > gpurun_out/r2_bench_c5_2gpu_final.json 2> gpurun_out/r2_bench_c5_2gpu_final.err; echo "c5 2gpu rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c5_2gpu_final.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"], "clk", j["clocks"]["sm_mhz"], "parity", j["parity"]["idx_equal"], j["parity"]["dist_bits_equal"], j["parity"]["topk"]["idx_equal"])
PY
