#!/bin/bash
# round-2 GPU call 11 (2 GPUs): single-process DEM over shards, C++ adapters with fir::n_gpus() = all devices
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_c1_shape.py -m gpu -x -q -rs 2>&1 | tail -8 | tee gpurun_out/r2_multigpu_tests_c.log
