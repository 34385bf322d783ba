#!/bin/bash
# BPTT kernel check: tests, in-kernel timelines, step trace and LSTM bench line per variant.  tools/gpu_bptt.sh TAG
mkdir -p gpurun_out
TAG=${1:-bptt}
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -m gpu tests/test_gpu_tc.py tests/test_gpu_base.py tests/test_gpu_bench_shapes.py -k "rnn_seq or train or lstm_bf16 or forward_backward or graph" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head
for ks in 16388 32772; do   # 4 | 64 << 8, 4 | 128 << 8
  echo "== SHOWTELL_BWD_KS=$ks"
  ST_BWD_KS=$ks timeout 120 python tools/timeline.py 2>&1 | grep -A5 "BWD per step" | tee gpurun_out/${TAG}_tl_$ks.txt
  SHOWTELL_BWD_KS=$ks timeout 600 python bench.py --no-extras --no-gpu-reference --no-cpu-baseline > gpurun_out/${TAG}_bench_$ks.json 2> gpurun_out/${TAG}_bench_$ks.err; echo "bench exit $?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_$ks.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print(d["roofline"].get("kernels_ms_per_step"))
PY
done
timeout 200 python tools/trace_step.py lstm 256 0 bf16 > gpurun_out/${TAG}_trace.txt 2>&1; grep -v Warn gpurun_out/${TAG}_trace.txt | sed -n 20,60p
