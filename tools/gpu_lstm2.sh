#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-l2}
timeout 200 python tools/trace_step.py lstm 256 0 bf16 > gpurun_out/${TAG}_trace.txt 2>&1; grep -v Warn gpurun_out/${TAG}_trace.txt | sed -n 20,60p
timeout 600 python bench.py --no-extras --no-gpu-reference --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print(d["roofline"].get("kernels_ms_per_step"))
PY
