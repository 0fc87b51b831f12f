#!/bin/bash
# round-2 GPU call 9 (1 GPU): suite + approximate chi2 / KL (entropy form) timings
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for m in chi2 kl; do timeout 150 python profiles/prof_approx.py $m 200000 1024 1280 10 2>&1 | tail -1 | tee gpurun_out/r2_prof_approx_${m}_v2.json; done
timeout 200 python profiles/prof_approx.py kl 1000000 1024 1280 10 2>&1 | tail -1 | tee gpurun_out/r2_prof_approx_kl_1M_v2.json
