#!/bin/bash
# round-2 GPU call 34 (1 GPU): folded norms (row norms in the spare columns of the shadow, raw-accumulator epilogue) — tests, soak, C4 / C2 lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_classifier_dem.py tests/test_gpu_property.py tests/test_gpu_fullsize.py tests/test_gpu_persist.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python -m pytest tests/test_gpu_fullsize_c345.py -m gpu -x -q -k c4 2>&1 | tail -2
timeout 300 python profiles/soak_large.py 40 11 > gpurun_out/r2_soak_large_f.jsonl 2>&1; echo "soak rc=$?"; tail -1 gpurun_out/r2_soak_large_f.jsonl; grep false gpurun_out/r2_soak_large_f.jsonl | cut -c1-300 | head -3
for c in c4 c2; do timeout 200 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu > gpurun_out/r2_bench_${c}_nf.json 2> gpurun_out/r2_bench_${c}_nf.err; echo "$c rc=$?"; python - <<PY
import json; j=json.loads(open("gpurun_out/r2_bench_${c}_nf.json").read().strip().splitlines()[-1]); print("ms/step", j["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms"], "parity", {k:v for k,v in j["parity"].items() if k.endswith("equal")})
PY
done
FIR_TENSOR_FOLD=0 timeout 200 python bench.py --config c4 --steps 10 --warmup 3 --skip-cpu --skip-parity 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nofold ms/step', j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'])"
