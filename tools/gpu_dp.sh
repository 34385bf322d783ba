#!/bin/bash
# data-parallel check on N GPUs: 2-GPU parity tests + the bench line (all workloads).  tools/gpu_dp.sh TAG N
mkdir -p gpurun_out
TAG=${1:-dp}; N=${2:-2}
if [ "$N" = "2" ]; then
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -m gpu tests/test_gpu_parallel.py > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|skipped" gpurun_out/${TAG}_pytest.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err; echo "bench exit $?"
tail -c 500 gpurun_out/${TAG}_n${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_n${N}.json").read().strip().splitlines()[-1])
def show(n, x):
    if "error" in x: print(n, x); return
    print(n, "| value", round(x["value"]), "| ms", round(x["ms_per_step"], 4), "| e2e", round(x["e2e"]["value"]), "| launches", x["gpu_launches"], "| n", x["n_gpus"])
show("main", d)
for n, x in d.get("others", {}).items(): show(n, x)
PY
