#!/bin/bash
# LSTM step check on the GPU box: base tests, in-kernel cluster timeline, step trace, bench line.  tools/gpu_lstm.sh TAG
mkdir -p gpurun_out
TAG=${1:-lstm}
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -m gpu tests/test_gpu_base.py tests/test_gpu_tc.py tests/test_gpu_bench_shapes.py -k "not beam and not greedy and not attention" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head
timeout 120 python tools/timeline_cluster.py 256 > gpurun_out/${TAG}_tl_cluster.txt 2>&1; tail -8 gpurun_out/${TAG}_tl_cluster.txt
timeout 200 python tools/trace_step.py lstm 256 0 bf16 > gpurun_out/${TAG}_trace.txt 2>&1; grep -v Warn gpurun_out/${TAG}_trace.txt | head -60
timeout 600 python bench.py --no-extras --no-gpu-reference --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print(d["roofline"].get("kernels_ms_per_step"))
PY
