"""Per-kernel device time of one eager training step measured by CUPTI (torch.profiler): warm caches, no replay."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, collections, re
import novel_view_synthesis_3d_b200 as P
from bench import make_host_batches
from torch.profiler import profile, ProfilerActivity
B, S = int(os.environ.get('XU_B', 8)), int(os.environ.get('XU_S', 64))
model = P.XUNet(dtype='bf16') if os.environ.get('XU_MODEL', 'small') == 'small' else P.XUNet.from_config(P.XUNetConfig(**{**P.FULL_3DIM.__dict__, 'dtype': 'bf16'}))
state = P.create_train_state(0, 1, 1e-4, B, S, model=model, init_on_device=os.environ.get('XU_MODEL', 'small') != 'small')
step = P.TrainStep(state, use_graph=False)
host = make_host_batches(2, B, S, 1234)
mask = np.ones(B, np.float32)
for i in range(5): step(host[i % 2][0], host[i % 2][1], cond_mask=mask)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3): step(host[i % 2][0], host[i % 2][1], cond_mask=mask)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = re.sub(r'\(anonymous namespace\)::', '', e.name)
        k = re.sub(r'\(.*', '', k)[:70]
        agg[k][0] += 1; agg[k][1] += e.device_time
tot = sum(v for _, v in agg.values())
print(f'total device us per step {tot/3:.1f}')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:44]:
    print(f'{v/tot*100:6.2f}% {v/3:9.1f}us/step n={c//3:4d} avg={v/c:7.2f}us  {k}')
if len(sys.argv) > 1:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and sys.argv[1] in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    print(sys.argv[1], 'per-launch us (one step, launch order):')
    print(' '.join(f'{e.device_time:.1f}' for e in evs[:n]))
