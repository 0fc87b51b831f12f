#!/bin/bash
# round-2 GPU call 13 (1 GPU): class-min on tensor cores with the cell-level overflow path + dense rerank list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_exact.py tests/test_gpu_property.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
timeout 120 python profiles/prof_classmin.py 100000 10000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_c2_tensor.json
timeout 200 python profiles/prof_classmin.py 2000000 20000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_2M_tensor.json
FIR_CLASSMIN_TENSOR=0 timeout 300 python profiles/prof_classmin.py 2000000 20000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_2M_exact.json
