#!/bin/bash
# Round-2 evidence run D (1 GPU): per-shape weight-gradient profile (one and two M tiles per CTA), fused-GroupNorm-backward K threshold,
# N-split rule on both workloads
set -u
O=gpurun_out
mkdir -p $O
XU_KERNEL=wgrad XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2m_wgrad_profile.txt 2>&1
XU_KERNEL=wgrad XUNET_WGRAD_M2=1 XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2m_wgrad_profile_m2.txt 2>&1
head -30 $O/r2m_wgrad_profile.txt
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
for v in default XUNET_GN_BWD_FUSED_MIN_K=4608 XUNET_CONV_BN_QUARTERS=4; do
  if [ "$v" = default ]; then e=""; else e="$v"; fi
  env $e timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2m_full_${v%%=*}.json 2> $O/r2m_full_${v%%=*}.err
done
for v in default XUNET_CONV_BN_QUARTERS=4; do
  if [ "$v" = default ]; then e=""; else e="$v"; fi
  env $e timeout 200 python bench.py --workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0 > $O/r2m_small_${v%%=*}.json 2> $O/r2m_small_${v%%=*}.err
done
for f in $O/r2m_full_*.json $O/r2m_small_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1); done
