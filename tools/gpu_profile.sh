#!/bin/bash
# GPU-box: full -m gpu test suite, bench, ncu launch list and one --set full capture of the fused denoiser kernel.
tag=${1:-r1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 46 2 > gpurun_out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:denoiser_tc -s 1 -c 1 -o gpurun_out/${tag}_tc python tools/profile_step.py 24 2 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
ls -la gpurun_out | tail -8
