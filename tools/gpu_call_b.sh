#!/bin/bash
# Round-2 evidence run B (1 GPU): where does the conv epilogue spend its time?  ncu source counters of the fused-statistics forward
# (EPI 1) and the FiLM GroupNorm-backward data gradient (EPI 3) at 128^2 / 256 channels; TMA-store epilogue forced everywhere;
# GPU suite + sanitizer on the fixed test cases.   Outputs -> gpurun_out/r2k_*
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2k_suite.log 2>&1
tail -3 $O/r2k_suite.log
SANITIZE_TIMEOUT=150 timeout 700 bash tools/sanitize.sh > $O/r2k_sanitize_summary.txt 2>&1
cat $O/r2k_sanitize_summary.txt
for t in memcheck racecheck synccheck; do gzip -f $O/sanitizer_$t.log; done
XUNET_CONV_TMA_STORE=1 XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2k_conv_step_profile_tmastore.txt 2>&1
head -12 $O/r2k_conv_step_profile_tmastore.txt
for epi in 1 3; do
  XU_MODEL=full XU_B=4 XU_S=128 timeout 600 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section LaunchStats \
    --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --import-source on --clock-control none --kernel-name-base demangled \
    -k regex:"conv_tc_kernel<\(int\)32, \(int\)${epi}>" -c 2 -o $O/r2k_epi${epi} -f python tools/run_step.py 1 > $O/r2k_ncu_epi${epi}.log 2>&1
  tail -2 $O/r2k_ncu_epi${epi}.log
  ncu -i $O/r2k_epi${epi}.ncu-rep --page raw --csv > $O/r2k_epi${epi}.csv 2>/dev/null
  ncu -i $O/r2k_epi${epi}.ncu-rep --page source --csv --print-source sass > $O/r2k_epi${epi}_sass.csv 2>/dev/null
  ncu -i $O/r2k_epi${epi}.ncu-rep --page source --csv --print-source cuda > $O/r2k_epi${epi}_cuda.csv 2>/dev/null
  gzip -f $O/r2k_epi${epi}_sass.csv $O/r2k_epi${epi}_cuda.csv
done
# wgrad with two M tiles per CTA (experimental): parity at the full widths, then the full-model step
XUNET_WGRAD_M2=1 timeout 240 python -m pytest tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider -k "conv_tcgen05_full_width" -s > $O/r2k_wgm2_optests.log 2>&1
tail -3 $O/r2k_wgm2_optests.log
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2k_full_default.json 2> $O/r2k_full_default.err
XUNET_WGRAD_M2=1 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2k_full_wgm2.json 2> $O/r2k_full_wgm2.err
XUNET_CONV_TMA_STORE=1 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2k_full_tmastore.json 2> $O/r2k_full_tmastore.err
grep -h -o '"ms_per_step": [0-9.]*' $O/r2k_full_default.json $O/r2k_full_wgm2.json $O/r2k_full_tmastore.json
ls -la $O/r2k_*
du -sh $O
