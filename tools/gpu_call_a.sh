#!/bin/bash
# Round-2 evidence run A (1 GPU): new 3-channel / weight-prep / adam / log-SNR kernels -- tests, A/B benches, CUPTI breakdown,
# ncu of the fused GroupNorm-backward dgrad epilogues at 128^2, sanitizer, launch list, batch sweep.  Outputs -> gpurun_out/r2j_*
set -u
O=gpurun_out
mkdir -p $O
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
# whole GPU suite first (stops at the first failure)
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2j_suite.log 2>&1
tail -3 $O/r2j_suite.log
# A/B on the full model (same box, same call)
for v in default XUNET_CONV3_OLD=1 XUNET_CONV_TMA_REDUCE=0 XUNET_CONV_BN_QUARTERS=4 XUNET_CONV_HALO_BK32=0; do
  if [ "$v" = default ]; then e=""; else e="$v"; fi
  env $e timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2j_full_${v%%=*}.json 2> $O/r2j_full_${v%%=*}.err
  case "$v" in default|XUNET_CONV3_OLD=1)
    env $e timeout 200 python bench.py --workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0 > $O/r2j_small_${v%%=*}.json 2> $O/r2j_small_${v%%=*}.err;;
  esac
done
grep -h -o '"ms_per_step": [0-9.]*' $O/r2j_full_*.json $O/r2j_small_*.json
# CUPTI per-kernel device time
XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/kineto_step.py > $O/r2j_kineto_full.txt 2>&1
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py > $O/r2j_kineto_small.txt 2>&1
# ncu: the GroupNorm-backward dgrad epilogues (EPI 3 = FiLM norm, EPI 2 = plain/swish norm) and a plain dgrad at 128^2, C=256
for epi in 3 0; do
  XU_MODEL=full XU_B=4 XU_S=128 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
    -k regex:"conv_tc_kernel<64, ${epi}>" --launch-skip $([ $epi = 0 ] && echo 60 || echo 0) -c 2 -o $O/r2j_epi${epi} -f python tools/run_step.py 1 > $O/r2j_ncu_epi${epi}.log 2>&1
  ncu -i $O/r2j_epi${epi}.ncu-rep --page raw --csv > $O/r2j_epi${epi}.csv 2>/dev/null
  ncu -i $O/r2j_epi${epi}.ncu-rep --page source --csv --print-source sass > $O/r2j_epi${epi}_sass.csv 2>/dev/null
  gzip -f $O/r2j_epi${epi}_sass.csv
done
ls -la $O/r2j_epi*
# sanitizer
SANITIZE_TIMEOUT=150 timeout 700 bash tools/sanitize.sh > $O/r2j_sanitize_summary.txt 2>&1
for t in memcheck racecheck synccheck; do gzip -f $O/sanitizer_$t.log; done
# ncu launch list of the bench command (bounded)
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2j_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-full128 --skip-cpu-baseline --sampler-steps 0 --no-graph > $O/r2j_launches.log 2>&1
gzip -f $O/r2j_launches.csv
# configs[4] batch sweep on one GPU
SWEEP_SHORT=1 timeout 600 bash tools/batch_sweep.sh 1 > $O/r2j_sweep_1gpu.jsonl 2> $O/r2j_sweep_1gpu.err
# pair-unit convolutions (experimental, off by default): op parity + model A/B + timing, each under its own timeout
XUNET_CONV_M2=1 timeout 240 python -m pytest tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider -k "conv_tcgen05_full_width" -s > $O/r2j_m2_optests.log 2>&1
tail -3 $O/r2j_m2_optests.log
timeout 240 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -p no:cacheprovider -k "pair_unit" -s > $O/r2j_m2_model.log 2>&1
tail -3 $O/r2j_m2_model.log
for bn in 128 256; do
  XUNET_CONV_M2=1 XUNET_CONV_M2_BN=$bn timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2j_full_m2_bn$bn.json 2> $O/r2j_full_m2_bn$bn.err
done
grep -h -o '"ms_per_step": [0-9.]*' $O/r2j_full_m2_bn*.json
XUNET_CONV_M2=1 XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2j_conv_step_profile_m2.txt 2>&1
XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2j_conv_step_profile.txt 2>&1
du -sh $O
