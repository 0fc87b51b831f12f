#!/bin/bash
# round-2 GPU call 19 (1 GPU): the whole GPU suite on the final tree, smoke(), latency-mode KL after the uniform-element step
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2_gputests_final.log 2>&1; echo "gpu suite rc=$?"; tail -22 gpurun_out/r2_gputests_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python profiles/prof_stream.py kl 1000000 1280 2>&1 | tail -4 | tee gpurun_out/r2_stream_kl_uniform.jsonl
