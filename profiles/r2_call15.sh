#!/bin/bash
# round-2 GPU call 15 (1 GPU): phased remainder + re-alignment (tests at small size, C5 A/B, ncu traffic), C2 regression check, C4 with the ratio sweep
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py tests/test_gpu_classifier_dem.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_1gpu_phased.json 2> gpurun_out/r2_bench_c5_1gpu_phased.err; echo "c5 phased rc=$?"; tail -c 2600 gpurun_out/r2_bench_c5_1gpu_phased.json
FIR_TENSOR_PHASED_MIN_BYTES=1000000000000 timeout 200 python bench.py --config c5 --steps 3 --warmup 3 --skip-parity --skip-cpu > gpurun_out/r2_bench_c5_1gpu_synconly.json 2> gpurun_out/r2_bench_c5_1gpu_synconly.err; echo "c5 sync-only rc=$?"; tail -c 1200 gpurun_out/r2_bench_c5_1gpu_synconly.json
timeout 200 python bench.py --config c2 --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_1gpu_b.json 2> gpurun_out/r2_bench_c2_1gpu_b.err; echo "c2 rc=$?"; tail -c 1500 gpurun_out/r2_bench_c2_1gpu_b.json
timeout 300 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r2_bench_c4_1gpu_sweep.json 2> gpurun_out/r2_bench_c4_1gpu_sweep.err; echo "c4 rc=$?"; tail -c 2500 gpurun_out/r2_bench_c4_1gpu_sweep.json; tail -3 gpurun_out/r2_bench_c4_1gpu_sweep.err
P5="python bench.py --config c5 --steps 1 --warmup 3 --skip-parity --skip-cpu"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:l2_candidates_kernel_2cta -s 4 -c 1 -f -o gpurun_out/r2_c5_candidates_phased $P5 > gpurun_out/r2_ncu_c5_phased_full.log 2>&1; echo "ncu c5 full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_c5_phased.csv $P5 > gpurun_out/r2_ncu_c5_phased_list.log 2>&1; echo "ncu c5 list rc=$?"
