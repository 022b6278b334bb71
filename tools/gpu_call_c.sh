#!/bin/bash
# Round-2 evidence run C (1 GPU): eight epilogue warps for the one-CTA-per-SM conv tiles -- parity, A/B, per-shape profile
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2l_suite.log 2>&1
tail -3 $O/r2l_suite.log
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2l_full_default.json 2> $O/r2l_full_default.err
XUNET_CONV_EW8=0 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2l_full_ew4.json 2> $O/r2l_full_ew4.err
XUNET_CONV_TMA_STORE=1 timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2l_full_tmastore.json 2> $O/r2l_full_tmastore.err
grep -h -o '"ms_per_step": [0-9.]*' $O/r2l_full_default.json $O/r2l_full_ew4.json $O/r2l_full_tmastore.json
XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2l_conv_step_profile.txt 2>&1
head -14 $O/r2l_conv_step_profile.txt
timeout 200 python bench.py --workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0 > $O/r2l_small_default.json 2> $O/r2l_small_default.err
grep -h -o '"ms_per_step": [0-9.]*' $O/r2l_small_default.json
