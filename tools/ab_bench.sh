#!/bin/bash
# GPU box: A/B of two builds of the library on the same GPU (alternating runs of the headline benchmark).
#   tools/ab_bench.sh <libA.so> <libB.so> [rounds]
A=$1; B=$2; R=${3:-2}
for i in $(seq $R); do
  for v in A B; do
    p=$A; [ $v = B ] && p=$B
    DFB200_LIB_PATH=$p python bench.py --suite headline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
