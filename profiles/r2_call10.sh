#!/bin/bash
# round-2 GPU call 10 (1 GPU): pipelined TMEM loads in the candidate epilogue — tensor tests + C2 / C4 / C5 lines, chi2 approx timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py tests/test_gpu_classifier_dem.py tests/test_gpu_property.py tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -4
for c in c2 c4; do timeout 120 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu > gpurun_out/r2_bench_${c}_d.json 2> gpurun_out/r2_bench_${c}_d.err; echo "$c rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_${c}_d.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "parity", {k:v for k,v in j["parity"].items() if k.endswith("equal")})
PY
done
timeout 200 python bench.py --config c5 --steps 4 --warmup 3 > gpurun_out/r2_bench_c5_d.json 2> gpurun_out/r2_bench_c5_d.err; echo "c5 rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c5_d.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"], "clk", j["clocks"]["sm_mhz"], "parity", {k:v for k,v in j["parity"].items() if k.endswith("equal")})
PY
timeout 150 python profiles/prof_approx.py chi2 200000 1024 1280 10 2>&1 | tail -1 | tee gpurun_out/r2_prof_approx_chi2_v3.json
