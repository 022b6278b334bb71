#!/bin/bash
# Round-2 evidence run I (1 GPU): (the two previous calls ran a stale library: a compile error had been hidden)  attention with deeper
# rings: parity + same-call A/B; wgrad 64-pixel ring A/B + per-shape profile
set -u
O=gpurun_out
mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider -k "attention" > $O/r2r_attn_tests.log 2>&1
tail -2 $O/r2r_attn_tests.log
grep -q "passed" $O/r2r_attn_tests.log || { echo "attention tests did not pass: stopping"; exit 0; }
grep -q "failed" $O/r2r_attn_tests.log && { echo "attention tests failed: stopping"; exit 0; }
PREV=novel_view_synthesis_3d_b200/libxunet_b200_prevattn.so
SM="--workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0"
timeout 200 python bench.py $SM > $O/r2r_small_new.json 2> $O/r2r_small_new.err
XUNET_LIB=$PREV timeout 200 python bench.py $SM > $O/r2r_small_prevattn.json 2> $O/r2r_small_prevattn.err
timeout 200 python bench.py $SM > $O/r2r_small_new2.json 2> $O/r2r_small_new2.err
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py > $O/r2r_kineto_small.txt 2>&1
head -5 $O/r2r_kineto_small.txt | tail -3
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2r_full_default.json 2> $O/r2r_full_default.err
XUNET_WGRAD_PT=64 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2r_full_wgpt64.json 2> $O/r2r_full_wgpt64.err
XUNET_WGRAD_M2=1 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2r_full_wgm2.json 2> $O/r2r_full_wgm2.err
for f in $O/r2r_small_*.json $O/r2r_full_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1); done
XU_KERNEL=wgrad XUNET_WGRAD_PT=64 XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2r_wgrad_profile_pt64.txt 2>&1
head -12 $O/r2r_wgrad_profile_pt64.txt
XUNET_WGRAD_PT=64 timeout 240 python -m pytest tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider -k "conv_tcgen05_full_width" > $O/r2r_wgpt64_optests.log 2>&1
tail -2 $O/r2r_wgpt64_optests.log
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2r_suite.log 2>&1
tail -3 $O/r2r_suite.log
