#!/bin/bash
# 2xB200: default bench line + data-parallel overlap experiments on the full model (bucket size, SM reserve, NCCL channel limits)
set -u
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --sampler-steps 0 > $O/r2v_2gpu_$tag.json 2> $O/r2v_2gpu_$tag.err
  python - $O/r2v_2gpu_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1]); f=d['full128']; a=f.get('allreduce',{})
    print(sys.argv[2], 'small ms', round(d['ms_per_step'],3), 'full ms', round(f['ms_per_step'],2), 'noar', round(a.get('step_ms_without_allreduce',0),2), 'alone', round(a.get('alone_ms',0),2), 'exposed', round(a.get('exposed_ms',0),2), 'buckets', a.get('buckets'))
except Exception as e: print(sys.argv[2],'ERR',e)
PY
}
run default A=1
run bucket64 XUNET_DP_BUCKET_MB=64
run bucket32 XUNET_DP_BUCKET_MB=32
run reserve8 XUNET_SM_RESERVE=8
run nchan8 NCCL_MAX_NCHANNELS=8
run nchan4_reserve4 NCCL_MAX_NCHANNELS=4 XUNET_SM_RESERVE=4
