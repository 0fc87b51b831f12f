#!/bin/bash
# round-2 GPU call 26 (1 GPU): soak — 200 freshly drawn problems per property test on the final tree
mkdir -p gpurun_out
FIR_PROPERTY_EXAMPLES=200 timeout 1500 python -m pytest tests/test_gpu_property.py -m gpu -q --durations=8 > gpurun_out/r2_property_soak.log 2>&1; echo "soak rc=$?"; tail -25 gpurun_out/r2_property_soak.log | cut -c1-400
