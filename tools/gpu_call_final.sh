#!/bin/bash
# final check of the round-2 tree (1 GPU): GPU suite, smoke(), default bench line
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3b_suite.log 2>&1
tail -2 $O/r3b_suite.log
true
true
timeout 600 python bench.py > $O/r3b_bench.json 2> $O/r3b_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3b_bench.json').read().splitlines() if l.startswith('{')][-1])
f=d['full128']
print('small', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'launches', d['kernels_per_step'], 'roofline', d['roofline']['kernel'], d['roofline']['per_launch_us'], d['roofline']['frac'])
print('full', f['ms_per_step'], f['images_per_sec'], f['step_tensor_frac_of_sustained'], [ (r['views_in_flight'], r['views_per_sec']) for r in f['sampler']['runs']])
print('sampler small', [(r['views_in_flight'], r['views_per_sec']) for r in d['sampler']['runs']], 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
