#!/bin/bash
# round-2 GPU call 32 (1 GPU): bench lines with the DRAM floor of the launch (roofline.traffic_floor)
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default_final2.json 2> gpurun_out/r2_bench_default_final2.err; echo "bench rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_default_final2.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], {k: j["roofline"].get(k) for k in ("frac","traffic","traffic_floor","traffic_floor_is")}, j["parity"]["idx_equal"], j["parity"]["topk"]["idx_equal"], j["cpu_baseline"]["value"])
PY
timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --skip-cpu > gpurun_out/r2_bench_c2_h.json 2> gpurun_out/r2_bench_c2_h.err; echo "c2 rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_c2_h.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], {k: j["roofline"].get(k) for k in ("frac","traffic","traffic_floor","traffic_floor_is")}, j["parity"]["idx_equal"])
PY
tail -2 gpurun_out/r2_bench_default_final2.err gpurun_out/r2_bench_c2_h.err 2>/dev/null | head
