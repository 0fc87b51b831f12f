#!/bin/bash
# round-2 GPU call 8 (1 GPU): full GPU suite + cls / c3-kl / c2 / c4 bench lines
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for c in cls c3-kl c2 c4; do
  timeout 200 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_${c}_c.json 2> gpurun_out/r2_bench_${c}_c.err; echo "$c rc=$?"
  tail -c 1500 gpurun_out/r2_bench_${c}_c.json; tail -3 gpurun_out/r2_bench_${c}_c.err
done
