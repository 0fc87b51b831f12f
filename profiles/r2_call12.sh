#!/bin/bash
# round-2 GPU call 12 (1 GPU): per-class nearest neighbour on tensor cores — tests + timings against the exact tiles; reference arm at C5
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_exact.py tests/test_gpu_property.py -m gpu -x -q 2>&1 | tail -15
timeout 120 python profiles/prof_classmin.py 100000 10000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_c2_tensor.json
FIR_CLASSMIN_TENSOR=0 timeout 120 python profiles/prof_classmin.py 100000 10000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_c2_exact.json
timeout 200 python profiles/prof_classmin.py 2000000 20000 512 1000 2>&1 | tail -1 | tee gpurun_out/r2_prof_classmin_2M_tensor.json
(time timeout 400 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2_bench_c5_reference.json 2> gpurun_out/r2_bench_c5_reference.err; echo "ref rc=$?"; tail -c 900 gpurun_out/r2_bench_c5_reference.json; tail -4 gpurun_out/r2_bench_c5_reference.err
