#!/bin/bash
# round-2 GPU call 29 (1 GPU): FFMA2 in the candidate epilogues — tests, soak, C4 / C2 / C5 lines, class-min timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py tests/test_gpu_classifier_dem.py tests/test_gpu_property.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/soak_large.py 30 7 > gpurun_out/r2_soak_large_e.jsonl 2>&1; echo "soak rc=$?"; tail -1 gpurun_out/r2_soak_large_e.jsonl
for c in c4 c2; do timeout 200 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu > gpurun_out/r2_bench_${c}_g.json 2> gpurun_out/r2_bench_${c}_g.err; echo "$c rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_${c}_g.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "parity", {k:v for k,v in j["parity"].items() if k.endswith("equal")}, j.get("k1"))
PY
done
timeout 300 python bench.py --config c5 --steps 4 --warmup 3 --skip-parity --skip-cpu > gpurun_out/r2_bench_c5_g.json 2> gpurun_out/r2_bench_c5_g.err; echo "c5 rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c5_g.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"], "clk", j["clocks"]["sm_mhz"], j.get("k1"))
PY
timeout 120 python profiles/prof_classmin.py 2000000 20000 512 1000 2>&1 | tail -1
