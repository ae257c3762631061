#!/bin/bash
# GPU-box dev cycle for the PointNet++ ops: parity tests, then the op-level benchmark (optionally with the
# ordered-scan ball query forced for an A/B line).  usage: tools/gpu_ops_cycle.sh <tag> [bench filters...]
tag=${1:-ops}; shift
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_pointnet2.py -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
timeout 300 python tools/bench_ops.py "$@" > gpurun_out/${tag}_ops.jsonl 2> gpurun_out/${tag}_ops.err
cut -c1-200 gpurun_out/${tag}_ops.jsonl; tail -3 gpurun_out/${tag}_ops.err
DFB200_BALL_QUERY=scan timeout 300 python tools/bench_ops.py ball_query > gpurun_out/${tag}_ops_scan.jsonl 2>> gpurun_out/${tag}_ops.err
cut -c1-200 gpurun_out/${tag}_ops_scan.jsonl
