#!/bin/bash
# round-2 GPU call 6 (2 GPUs): sharded DEM + NCCL paths, the corrected tests, KL kernel timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_exact.py tests/test_gpu_classifier_dem.py tests/test_gpu_c1_shape.py -m gpu -x -q -rs 2>&1 | tail -25 | tee gpurun_out/r2_multigpu_tests_b.log
timeout 120 python profiles/prof_exact.py kl 200000 512 1280 2>&1 | tail -4 | tee gpurun_out/r2_prof_exact_kl_rolled.log
