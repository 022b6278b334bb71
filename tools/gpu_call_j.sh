#!/bin/bash
# Round-2 evidence run J (1 GPU): wgrad with a second TMA producer lane; fused attention backward with per-warp constants
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider > $O/r2s_op_tests.log 2>&1
tail -2 $O/r2s_op_tests.log
grep -q "passed" $O/r2s_op_tests.log || { echo "op tests did not pass: stopping"; exit 0; }
grep -q "failed" $O/r2s_op_tests.log && { echo "op tests failed: stopping"; exit 0; }
PREV=novel_view_synthesis_3d_b200/libxunet_b200_prev2.so
SM="--workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0"
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
timeout 200 python bench.py $SM > $O/r2s_small_new.json 2> $O/r2s_small_new.err
XUNET_LIB=$PREV timeout 200 python bench.py $SM > $O/r2s_small_prev.json 2> $O/r2s_small_prev.err
timeout 200 python bench.py $SM > $O/r2s_small_new2.json 2> $O/r2s_small_new2.err
timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2s_full_new.json 2> $O/r2s_full_new.err
XUNET_WGRAD_TWO_PRODUCERS=0 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2s_full_oneprod.json 2> $O/r2s_full_oneprod.err
XUNET_LIB=$PREV timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2s_full_prev.json 2> $O/r2s_full_prev.err
for f in $O/r2s_small_*.json $O/r2s_full_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1); done
XU_KERNEL=wgrad XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2s_wgrad_profile.txt 2>&1
head -14 $O/r2s_wgrad_profile.txt
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py > $O/r2s_kineto_small.txt 2>&1
head -5 $O/r2s_kineto_small.txt | tail -3
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2s_suite.log 2>&1
tail -3 $O/r2s_suite.log
