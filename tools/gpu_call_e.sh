#!/bin/bash
# Round-2 evidence run E (1 GPU): attention backward with pre-scaled per-query constants -- parity + same-call A/B against the
# previous build of attn_tc.cu (XUNET_LIB), N-split rule by reduction length on the small model
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2n_suite.log 2>&1
tail -3 $O/r2n_suite.log
PREV=novel_view_synthesis_3d_b200/libxunet_b200_prevattn.so
SM="--workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0"
timeout 200 python bench.py $SM > $O/r2n_small_new.json 2> $O/r2n_small_new.err
XUNET_LIB=$PREV timeout 200 python bench.py $SM > $O/r2n_small_prevattn.json 2> $O/r2n_small_prevattn.err
timeout 200 python bench.py $SM > $O/r2n_small_new2.json 2> $O/r2n_small_new2.err
XUNET_CONV_BN_QUARTERS=4 timeout 200 python bench.py $SM > $O/r2n_small_bnq4.json 2> $O/r2n_small_bnq4.err
BQ="--steps 8 --warmup 3 --skip-cpu-baseline --sampler-steps 0"
timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2n_full_new.json 2> $O/r2n_full_new.err
XUNET_LIB=$PREV timeout 300 python bench.py --workload full128 --batch 4 $BQ > $O/r2n_full_prevattn.json 2> $O/r2n_full_prevattn.err
for f in $O/r2n_small_*.json $O/r2n_full_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1) $(python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
r=d.get('roofline') or {}
print(r.get('kernel'), round(r.get('per_launch_us',0),2), 'us frac', round(r.get('frac',0),4))
PY
); done
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py > $O/r2n_kineto_small.txt 2>&1
head -12 $O/r2n_kineto_small.txt
