#!/bin/bash
# round-2 GPU call 18 (1 GPU): at-size tests C3 / C4 / C5 after the latency-mode fix, streaming tests, ncu of the compact KL tile kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize_c345.py -m gpu -x -q --durations=6 > gpurun_out/r2_test_c345.log 2>&1; echo "c345 rc=$?"; tail -12 gpurun_out/r2_test_c345.log
PK="python profiles/prof_exact.py kl 200000 512 1280"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:exact_tile_kernel -s 1 -c 1 -f -o gpurun_out/r2_exact_kl_compact $PK > gpurun_out/r2_ncu_exact_kl_compact.log 2>&1; echo "ncu kl rc=$?"
timeout 200 python bench.py --config c3-chi2 --steps 2 --warmup 3 > gpurun_out/r2_bench_c3-chi2_1gpu_b.json 2> gpurun_out/r2_bench_c3-chi2_1gpu_b.err; echo "c3-chi2 rc=$?"; tail -c 600 gpurun_out/r2_bench_c3-chi2_1gpu_b.json
