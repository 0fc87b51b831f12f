#!/bin/bash
# round-2 GPU call 16 (1 GPU): KL tile kernel with the warp-uniform query element (zero elements skip the logf pair), at-size parity tests C3 / C4 / C5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_property.py tests/test_gpu_golden.py tests/test_gpu_twd.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_fullsize_c345.py -m gpu -x -q --durations=6 2>&1 | tail -14
timeout 300 python bench.py --config c3-kl --steps 2 --warmup 3 > gpurun_out/r2_bench_c3-kl_1gpu_uniform.json 2> gpurun_out/r2_bench_c3-kl_1gpu_uniform.err; echo "c3-kl rc=$?"; tail -c 2200 gpurun_out/r2_bench_c3-kl_1gpu_uniform.json
timeout 120 python profiles/prof_exact.py kl 200000 512 1280 2>&1 | tail -2
