#!/bin/bash
# usage: tools/ab_bench.sh <tag> "ENV=1 ..." ["ENV2=..."]...  — bench.py (no extras, no cpu) once per environment setting; prints ms by kernel
tag=$1; shift
for envs in "$@"; do
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$envs', 'step ms', round(d['ms_per_step'], 2), d['roofline']['ms_by_kernel'])" | tee -a gpurun_out/ab_$tag.txt
done
