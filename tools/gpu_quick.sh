#!/bin/bash
# quick GPU check of selected tests:  tools/gpu_quick.sh TAG <pytest args...>
mkdir -p gpurun_out
TAG=$1; shift
timeout 1200 python -m pytest -q -s --tb=short -p no:cacheprovider -m gpu "$@" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"
grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head -40
