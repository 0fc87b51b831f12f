#!/bin/bash
# round-2 GPU call 27 (1 GPU): large-shape soak of the fast paths against the exact kernels
mkdir -p gpurun_out
timeout 900 python profiles/soak_large.py 14 1 > gpurun_out/r2_soak_large.jsonl 2> gpurun_out/r2_soak_large.err; echo "soak rc=$?"; cut -c1-420 gpurun_out/r2_soak_large.jsonl; tail -5 gpurun_out/r2_soak_large.err
