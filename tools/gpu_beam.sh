#!/bin/bash
# decode-path check on the GPU box: decode tests + the beam-3 bench line.  tools/gpu_beam.sh TAG
mkdir -p gpurun_out
TAG=${1:-beam}
timeout 1200 python -m pytest -q -s --tb=short -p no:cacheprovider -m gpu tests/test_gpu_tc.py tests/test_gpu_base.py tests/test_gpu_bench_shapes.py -k "topk or greedy or beam or decode" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"
grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -3
grep -E "^FAILED|^ERROR|identical|separated" gpurun_out/${TAG}_pytest.log | head -60
for wl in beam3 beam5; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "bench $wl exit $?"
tail -c 400 gpurun_out/${TAG}_$wl.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_$wl.json").read().strip().splitlines()[-1])
    print("$wl value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"], 4))
    print(d["roofline"].get("kernels_ms_per_step"))
except Exception as e:
    print("parse failed", e)
PY
done
