#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?"
grep -E "inside=|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -n 40
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -n 5 gpurun_out/smoke.log
