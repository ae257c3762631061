#!/bin/bash
# GPU-box: round-end check list -- full -m gpu suite, smoke, training-step benchmark, op-level benchmark, headline bench.
tag=${1:-final}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_tests.log 2>&1; cat gpurun_out/${tag}_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 100 python tools/bench_train.py 16 bf16 2>/dev/null | tee gpurun_out/${tag}_train.jsonl
timeout 100 python tools/bench_train.py 16 fp32 2>/dev/null | tee -a gpurun_out/${tag}_train.jsonl
timeout 300 python tools/bench_ops.py > gpurun_out/${tag}_ops.jsonl 2> gpurun_out/${tag}_ops.err; wc -l gpurun_out/${tag}_ops.jsonl
timeout 250 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json
