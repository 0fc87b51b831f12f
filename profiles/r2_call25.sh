#!/bin/bash
# round-2 GPU call 25 (1 GPU): final tree — whole GPU suite, smoke, the default bench as the driver runs it; prefix brute force on the tensor path at 100k x 10k;
# ncu of the class-min passes
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_gputests_final.log 2>&1; echo "gpu suite rc=$?"; tail -14 gpurun_out/r2_gputests_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; echo "bench rc=$?"; tail -c 2800 gpurun_out/r2_bench_default_final.json; tail -4 gpurun_out/r2_bench_default_final.err
timeout 200 python profiles/prof_twd.py 100000 10000 512 1000 4 2>&1 | tail -3 | cut -c1-1500 | tee gpurun_out/r2_twd_100k.jsonl
PC="python profiles/prof_classmin.py 2000000 20000 512 1000"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:l2_candidates_kernel_2cta -s 2 -c 2 -f -o gpurun_out/r2_classmin_passes $PC > gpurun_out/r2_ncu_classmin.log 2>&1; echo "ncu classmin rc=$?"
