#!/bin/bash
# GPU-box dev cycle: denoiser parity tests, bench, steady-state timeline.  usage: tools/gpu_cycle.sh <tag>
tag=${1:-dev}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_denoiser.py tests/test_gpu_umma.py -x -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json | cut -c1-260; tail -3 gpurun_out/${tag}_bench.err
timeout 100 python tools/diag_timeline.py 3 > gpurun_out/${tag}_timeline.txt 2>&1
cat gpurun_out/${tag}_timeline.txt
