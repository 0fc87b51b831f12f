#!/bin/bash
# round-2 GPU call 1: sanity of the shipped tree + pipe peaks + ncu evidence for the chi2/KL kernels
mkdir -p gpurun_out
profiles/peak_pipes > gpurun_out/r2_peak_pipes.json 2> gpurun_out/r2_peak_pipes.err
cat gpurun_out/r2_peak_pipes.json
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in chi2 kl; do
  timeout 300 python profiles/prof_exact.py $m 200000 512 1280 > gpurun_out/r2_prof_exact_$m.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_tile_kernel -s 1 -c 1 -f -o gpurun_out/r2_exact_$m \
      python profiles/prof_exact.py $m 200000 512 1280 > gpurun_out/r2_ncu_exact_$m.log 2>&1
  timeout 300 python profiles/prof_approx.py $m 200000 1024 1280 10 > gpurun_out/r2_prof_approx_$m.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:approx_tile_kernel -s 1 -c 1 -f -o gpurun_out/r2_approx_$m \
      python profiles/prof_approx.py $m 200000 1024 1280 10 > gpurun_out/r2_ncu_approx_$m.log 2>&1
done
tail -2 gpurun_out/r2_prof_exact_*.log gpurun_out/r2_prof_approx_*.log
ls -la gpurun_out/*.ncu-rep
