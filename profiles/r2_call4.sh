#!/bin/bash
# round-2 GPU call 4 (1 GPU): full GPU test suite with the tensor-core DEM first round + the secondary bench configs
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for c in c4 c3-chi2 c3-kl c1; do
  timeout 170 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2_bench_${c}_a.json 2> gpurun_out/r2_bench_${c}_a.err; echo "$c rc=$?"
  tail -c 2800 gpurun_out/r2_bench_${c}_a.json; tail -4 gpurun_out/r2_bench_${c}_a.err
done
