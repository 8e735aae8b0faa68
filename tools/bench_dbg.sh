#!/bin/bash
# usage: tools/bench_dbg.sh <cfg> <dbg...>   — prints conv_tc ms per step for each debug mask
c=$1; shift
for d in "$@"; do
  B200AD_CONV_CFG=$c B200AD_CONV_DBG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('cfg', $c, 'dbg', $d, 'conv_tc ms', d['roofline']['ms_by_kernel']['conv_tc'], 'step ms', round(d['ms_per_step'], 2))"
done
