#!/bin/bash
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
