#!/bin/bash
# round-2 GPU call 24 (1 GPU): pair-list rerank strides over the device-side count (grid sized for the machine) — tests, C5 / C2 phases, class-min timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py tests/test_gpu_property.py tests/test_gpu_exact.py tests/test_gpu_classifier_dem.py -m gpu -x -q 2>&1 | tail -4
timeout 100 python profiles/prof_phases.py 10 10000000 100000 2 2>&1 | tail -1 | tee gpurun_out/r2_phases_c5_k10_trim2.json
timeout 100 python profiles/prof_phases.py 10 100000 10000 20 2>&1 | tail -1 | tee gpurun_out/r2_phases_c2_k10_trim2.json
timeout 120 python profiles/prof_classmin.py 100000 10000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_c2_tensor_b.json
timeout 150 python profiles/prof_approx.py kl 200000 1024 1280 10 2>&1 | tail -1 | tee gpurun_out/r2_prof_approx_kl_b.json
