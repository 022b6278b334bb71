#!/bin/bash
# Round-2 evidence run F (1 GPU): software-pipelined fused attention backward -- parity (under short timeouts: new barrier protocol),
# then same-call A/B against the previous build of attn_tc.cu
set -u
O=gpurun_out
mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attention" > $O/r2p_attn_tests.log 2>&1
tail -3 $O/r2p_attn_tests.log
grep -q "failed\|Error\|Timeout" $O/r2p_attn_tests.log && { echo "attention tests failed: stopping"; exit 0; }
grep -q "passed" $O/r2p_attn_tests.log || { echo "attention tests did not finish: stopping"; exit 0; }
PREV=novel_view_synthesis_3d_b200/libxunet_b200_prevattn.so
SM="--workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0"
timeout 200 python bench.py $SM > $O/r2p_small_new.json 2> $O/r2p_small_new.err
XUNET_LIB=$PREV timeout 200 python bench.py $SM > $O/r2p_small_prevattn.json 2> $O/r2p_small_prevattn.err
timeout 200 python bench.py $SM > $O/r2p_small_new2.json 2> $O/r2p_small_new2.err
for f in $O/r2p_small_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1) $(python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
r=d.get('roofline') or {}
print(r.get('kernel'), round(r.get('per_launch_us',0),2), 'us frac', round(r.get('frac',0),4))
PY
); done
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py > $O/r2p_kineto_small.txt 2>&1
head -8 $O/r2p_kineto_small.txt | tail -6
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2p_suite.log 2>&1
tail -3 $O/r2p_suite.log
