#!/bin/bash
# compute-sanitizer evidence: memcheck + racecheck + synccheck over the smoke invocation (forward + fused
# scheduler step: conv_tc / conv_in / conv_out / attention / temb kernels) and over one small backward
# (unet_bwd / wgrad_tc / bwd kernels).  Logs go to gpurun_out/sanitizer_<tool>_<what>.log; copy the summaries to profiles/.
# usage (on the GPU box): tools/sanitize.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_${tool}_smoke_${tag}.log \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_smoke_${tag}.out 2>&1
  echo "$tool smoke: exit $?"
  tail -3 gpurun_out/sanitizer_${tool}_smoke_${tag}.log
done
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_${tool}_bwd_${tag}.log \
    python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "test_unet_backward_matches_autograd" \
    > gpurun_out/sanitizer_${tool}_bwd_${tag}.out 2>&1
  echo "$tool backward: exit $?"
  tail -3 gpurun_out/sanitizer_${tool}_bwd_${tag}.log
done
