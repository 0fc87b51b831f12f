#!/bin/bash
# round-2 GPU call 28 (1 GPU): longer large-shape soaks (default switches; forced phased remainder + re-alignment every 8 tiles; single-CTA kernel)
mkdir -p gpurun_out
timeout 900 python profiles/soak_large.py 70 2 > gpurun_out/r2_soak_large_b.jsonl 2> gpurun_out/r2_soak_large_b.err; echo "soak b rc=$?"; tail -1 gpurun_out/r2_soak_large_b.jsonl; grep -c '"topk": true' gpurun_out/r2_soak_large_b.jsonl; grep false gpurun_out/r2_soak_large_b.jsonl | cut -c1-400 | head -5
FIR_TENSOR_PHASED_MIN_BYTES=0 FIR_TENSOR_SYNC_TILES=8 timeout 900 python profiles/soak_large.py 40 3 > gpurun_out/r2_soak_large_c.jsonl 2> gpurun_out/r2_soak_large_c.err; echo "soak c rc=$?"; tail -1 gpurun_out/r2_soak_large_c.jsonl; grep false gpurun_out/r2_soak_large_c.jsonl | cut -c1-400 | head -5
FIR_TENSOR_SEED=0 FIR_TENSOR_R1=4 timeout 900 python profiles/soak_large.py 25 4 > gpurun_out/r2_soak_large_d.jsonl 2> gpurun_out/r2_soak_large_d.err; echo "soak d rc=$?"; tail -1 gpurun_out/r2_soak_large_d.jsonl; grep false gpurun_out/r2_soak_large_d.jsonl | cut -c1-400 | head -5
tail -3 gpurun_out/r2_soak_large_b.err gpurun_out/r2_soak_large_c.err gpurun_out/r2_soak_large_d.err
