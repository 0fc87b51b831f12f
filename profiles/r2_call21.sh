#!/bin/bash
# round-2 GPU call 21 (1 GPU): prefix distances on the tensor path, sqrt(N) seed sample; regression lines C2 / C4 / C5; latency-mode KL (branch-free form)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_exact.py tests/test_gpu_twd.py tests/test_gpu_compat_cpp.py tests/test_gpu_classifier_dem.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
for c in c2 c4; do timeout 200 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu > gpurun_out/r2_bench_${c}_e.json 2> gpurun_out/r2_bench_${c}_e.err; echo "$c rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_${c}_e.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "parity", {k:v for k,v in j["parity"].items() if k.endswith("equal")}, j.get("k1"))
PY
done
timeout 300 python bench.py --config c5 --steps 4 --warmup 3 --skip-parity --skip-cpu > gpurun_out/r2_bench_c5_e.json 2> gpurun_out/r2_bench_c5_e.err; echo "c5 rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c5_e.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"], "clk", j["clocks"]["sm_mhz"], j.get("k1"), j.get("certificate_fallback_queries"))
PY
timeout 100 python profiles/prof_phases.py 10 10000000 100000 2 2>&1 | tail -1 | tee gpurun_out/r2_phases_c5_k10_final.json
timeout 200 python profiles/prof_stream.py kl 1000000 1280 2>&1 | tail -4 | tee gpurun_out/r2_stream_kl_branchfree.jsonl
timeout 100 python profiles/prof_twd.py 2>&1 | tail -3
