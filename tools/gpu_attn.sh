#!/bin/bash
# attention decoders on the GPU box: tests, step trace, bench line.  tools/gpu_attn.sh TAG
mkdir -p gpurun_out
TAG=${1:-attn}
timeout 1200 python -m pytest -q --tb=short -p no:cacheprovider -m gpu tests/test_gpu_attn.py tests/test_gpu_kernels.py tests/test_gpu_bench_shapes.py -k "not beam and not lstm_bf16" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head
timeout 200 python tools/trace_step.py attn_gru 128 196 bf16 > gpurun_out/${TAG}_trace.txt 2>&1; grep -v Warn gpurun_out/${TAG}_trace.txt | sed -n 1,32p
timeout 600 python bench.py --workload attn_gru_train --no-gpu-reference --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print(d["roofline"].get("kernels_ms_per_step"))
PY
