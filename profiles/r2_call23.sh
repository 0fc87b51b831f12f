#!/bin/bash
# round-2 GPU call 23 (1 GPU): prune / select scan only the lists a query block owns — tests + C5 phases + C2 line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py tests/test_gpu_property.py tests/test_gpu_classifier_dem.py tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_fullsize_c345.py -m gpu -x -q -k "c5 or c4" 2>&1 | tail -3
timeout 100 python profiles/prof_phases.py 10 10000000 100000 2 2>&1 | tail -1 | tee gpurun_out/r2_phases_c5_k10_trim.json
timeout 100 python profiles/prof_phases.py 10 100000 10000 20 2>&1 | tail -1 | tee gpurun_out/r2_phases_c2_k10_trim.json
timeout 200 python bench.py --config c2 --steps 20 --warmup 3 --skip-cpu > gpurun_out/r2_bench_c2_f.json 2> gpurun_out/r2_bench_c2_f.err; echo "c2 rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c2_f.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "parity", {k:v for k,v in j["parity"].items() if k.endswith("equal")}, j.get("k1"))
PY
