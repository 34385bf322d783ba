#!/bin/bash
# One gpurun call: GPU tests (+ printed parity tables), smoke, the default bench line (with the other workloads and the
# reference bars) and the reference arm.  Output -> gpurun_out/.   tools/r02_check.sh TAG [pytest args]
mkdir -p gpurun_out
TAG=${1:-r02a}; shift
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider "$@" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest.log | tail -5
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    def show(n, x):
        if "error" in x: print(n, x); return
        print(n, "| value", round(x["value"]), x["unit"], "| ms", round(x["ms_per_step"], 3), "| e2e", round(x["e2e"]["value"]), "| launches", x["gpu_launches"])
        print("   roofline:", {k: x["roofline"].get(k) for k in ("kernel", "achieved", "frac", "step_tflops", "step_frac_of_sustained_peak")})
        print("   kernels ms/step:", x["roofline"]["kernels_ms_per_step"])
        print("   gpu_reference:", x.get("gpu_reference")); print("   cpu_baseline:", x.get("cpu_baseline"))
    show("main", d)
    for n, x in d.get("others", {}).items(): show(n, x)
except Exception as e:
    print("bench parse failed", e)
PY
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref exit $?"
cut -c1-400 gpurun_out/${TAG}_ref.json
