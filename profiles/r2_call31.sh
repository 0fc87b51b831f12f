#!/bin/bash
# round-2 GPU call 31 (8 GPUs): C5 over 8 GPUs on the final tree, parity on the merged result
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 \
  > gpurun_out/r2_bench_c5_8gpu_final.json 2> gpurun_out/r2_bench_c5_8gpu_final.err; echo "c5 8gpu rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c5_8gpu_final.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"], "clk", j["clocks"]["sm_mhz"], "parity", j["parity"]["idx_equal"], j["parity"]["dist_bits_equal"], j["parity"]["topk"]["idx_equal"])
PY
