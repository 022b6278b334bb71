#!/bin/bash
# compute-sanitizer evidence (run on a GPU box from the repo root): memcheck + racecheck + synccheck over (a) smoke() = one tiny
# forward + backward + Adam of the whole network in fp32 (SIMT kernels) and bf16 (tcgen05 kernels, fused epilogues) and (b) the
# operator-level tcgen05 tests at their smallest shapes.  Logs -> gpurun_out/sanitizer_<tool>.log (copy the summaries to profiles/).
set -u
mkdir -p gpurun_out
SEL='test_conv_tcgen05_fwd_and_dgrad and (case8 or case9) or test_attention_tcgen05_fwd and case0 or test_attention_tcgen05_bwd and case0 or test_fused_qkv or test_three_channel'
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitizer_${tool}.log
  echo "== compute-sanitizer --tool $tool : smoke()" > $log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke >> $log 2>&1
  echo "== compute-sanitizer --tool $tool : operator tests" >> $log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL" -p no:cacheprovider >> $log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke OK" $log | sed "s/^/[$tool] /"
done
