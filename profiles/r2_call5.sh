#!/bin/bash
# round-2 GPU call 5 (1 GPU): full test suite, C3-KL / C5 bench lines, ncu captures for C5 and C4 (each after the same command ran clean)
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 240 python bench.py --config c3-kl --steps 2 --warmup 3 > gpurun_out/r2_bench_c3-kl_b.json 2> gpurun_out/r2_bench_c3-kl_b.err; echo "c3-kl rc=$?"; tail -c 1500 gpurun_out/r2_bench_c3-kl_b.json
timeout 200 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_b.json 2> gpurun_out/r2_bench_c5_b.err; echo "c5 rc=$?"; tail -c 1800 gpurun_out/r2_bench_c5_b.json
P5="python bench.py --config c5 --steps 1 --warmup 3 --skip-parity --skip-cpu"
P4="python bench.py --config c4 --steps 1 --warmup 3 --skip-parity --skip-cpu"
timeout 120 $P5 > gpurun_out/r2_prof_c5_plain.json 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:l2_candidates_kernel_2cta -s 4 -c 1 -f -o gpurun_out/r2_c5_candidates $P5 > gpurun_out/r2_ncu_c5_full.log 2>&1; echo "ncu c5 full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_c5.csv $P5 > gpurun_out/r2_ncu_c5_list.log 2>&1; echo "ncu c5 list rc=$?"
timeout 120 $P4 > gpurun_out/r2_prof_c4_plain.json 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:l2_candidates_kernel_2cta -s 4 -c 1 -f -o gpurun_out/r2_c4_dem_candidates $P4 > gpurun_out/r2_ncu_c4_full.log 2>&1; echo "ncu c4 full rc=$?"
ls -la gpurun_out/*.ncu-rep
