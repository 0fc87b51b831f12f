#!/bin/bash
# round-2 GPU call 2: GPU tests of the new tree + first C5 / C2 bench lines (inner timeouts sum to < the gpurun limit)
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_a.json 2> gpurun_out/r2_bench_c5_a.err; echo "c5 rc=$?"; tail -c 3000 gpurun_out/r2_bench_c5_a.json; tail -5 gpurun_out/r2_bench_c5_a.err
timeout 100 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err; echo "c2 rc=$?"; tail -c 3000 gpurun_out/r2_bench_c2_a.json; tail -5 gpurun_out/r2_bench_c2_a.err
