#!/bin/bash
# round-2 GPU call 14 (1 GPU): pair re-alignment in the full rounds of the candidate kernel (C5 A/B + ncu traffic), k = 1 list length at C5
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_1gpu_sync.json 2> gpurun_out/r2_bench_c5_1gpu_sync.err; echo "c5 sync rc=$?"; tail -c 2500 gpurun_out/r2_bench_c5_1gpu_sync.json
FIR_TENSOR_SYNC=0 timeout 200 python bench.py --config c5 --steps 3 --warmup 3 --skip-parity --skip-cpu > gpurun_out/r2_bench_c5_1gpu_nosync.json 2> gpurun_out/r2_bench_c5_1gpu_nosync.err; echo "c5 nosync rc=$?"; tail -c 1200 gpurun_out/r2_bench_c5_1gpu_nosync.json
for r1 in 0 8 16; do
  FIR_TENSOR_R1=$r1 timeout 120 python profiles/prof_phases.py 1 10000000 100000 2 2>&1 | tail -1 | tee gpurun_out/r2_phases_c5_k1_R$r1.json
done
P5="python bench.py --config c5 --steps 1 --warmup 3 --skip-parity --skip-cpu"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:l2_candidates_kernel_2cta -s 4 -c 1 -f -o gpurun_out/r2_c5_candidates_sync $P5 > gpurun_out/r2_ncu_c5_sync_full.log 2>&1; echo "ncu c5 full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
