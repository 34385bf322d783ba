#!/bin/bash
# Dev loop on the GPU box: all gpu tests (no -x, short tracebacks) + timing. Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider "$@" > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
timeout 300 python tools/time_base.py > gpurun_out/time_base.log 2>&1
tail -45 gpurun_out/time_base.log
