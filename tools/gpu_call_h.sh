#!/bin/bash
# Round-2 evidence run H (1 GPU): ncu source counters of the attention kernels (small model), wgrad 64-pixel ring A/B
set -u
O=gpurun_out
mkdir -p $O
for k in attn_bwd_fused_tc_kernel attn_fwd_tc_kernel; do
  timeout 300 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section LaunchStats --section Occupancy \
    --section SchedulerStats --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:$k --launch-skip 12 -c 1 -o $O/r2q_$k -f python tools/run_step.py 2 > $O/r2q_ncu_$k.log 2>&1
  tail -1 $O/r2q_ncu_$k.log
  ncu -i $O/r2q_$k.ncu-rep --page raw --csv > $O/r2q_$k.csv 2>/dev/null
  ncu -i $O/r2q_$k.ncu-rep --page source --csv --print-source sass > $O/r2q_${k}_sass.csv 2>/dev/null
  gzip -f $O/r2q_${k}_sass.csv
done
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2q_full_default.json 2> $O/r2q_full_default.err
XUNET_WGRAD_PT=64 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2q_full_wgpt64.json 2> $O/r2q_full_wgpt64.err
for f in $O/r2q_full_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1); done
XU_KERNEL=wgrad XUNET_WGRAD_PT=64 XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2q_wgrad_profile_pt64.txt 2>&1
head -12 $O/r2q_wgrad_profile_pt64.txt
XUNET_WGRAD_PT=64 timeout 240 python -m pytest tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider -k "conv_tcgen05_full_width" > $O/r2q_wgpt64_optests.log 2>&1
tail -2 $O/r2q_wgpt64_optests.log
ls -la $O/r2q_*
