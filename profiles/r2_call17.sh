#!/bin/bash
# round-2 GPU call 17 (1 GPU): compact KL tile loop (accumulators in the output tile), finer splits for the class modes; at-size tests with full tracebacks
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize_c345.py -m gpu -x -q -k "kl" > gpurun_out/r2_test_c345_kl.log 2>&1; echo "kl alone rc=$?"; grep -E "Error|error|passed|failed" gpurun_out/r2_test_c345_kl.log | head -20
timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_property.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize_c345.py -m gpu -q --durations=6 > gpurun_out/r2_test_c345.log 2>&1; echo "c345 rc=$?"; tail -12 gpurun_out/r2_test_c345.log
timeout 300 python bench.py --config c3-kl --steps 2 --warmup 3 > gpurun_out/r2_bench_c3-kl_1gpu_compact.json 2> gpurun_out/r2_bench_c3-kl_1gpu_compact.err; echo "c3-kl rc=$?"; tail -c 1500 gpurun_out/r2_bench_c3-kl_1gpu_compact.json
timeout 120 python profiles/prof_exact.py kl 200000 512 1280 2>&1 | tail -2
PK="python bench.py --config c3-kl --steps 1 --warmup 3 --skip-parity --skip-cpu"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:exact_tile_kernel -s 1 -c 1 -f -o gpurun_out/r2_exact_kl_compact $PK > gpurun_out/r2_ncu_exact_kl_compact.log 2>&1; echo "ncu kl rc=$?"
