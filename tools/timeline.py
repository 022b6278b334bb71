"""Kernel timeline of one graph-replayed training step (CUPTI via torch.profiler): per-stream gaps and the critical path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, re
import novel_view_synthesis_3d_b200 as P
from bench import make_host_batches
from torch.profiler import profile, ProfilerActivity
B, S = 8, 64
model = P.XUNet(dtype='bf16')
state = P.create_train_state(0, 1, 1e-4, B, S, model=model)
step = P.TrainStep(state, use_graph=True)
host = make_host_batches(2, B, S, 1234)
mask = np.ones(B, np.float32)
for i in range(6): step(host[i % 2][0], host[i % 2][1], cond_mask=mask)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(host[0][0], host[0][1], cond_mask=mask)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
gap_min = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
last_end = {}
streams = {}
for e in evs:
    st = getattr(e, 'device_resource_id', None)
    if st is None: st = 0
    streams.setdefault(st, len(streams))
    name = re.sub(r'\(anonymous namespace\)::', '', e.name); name = re.sub(r'\(.*', '', name)[:48]
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = s - last_end.get(st, s)
    last_end[st] = s + d
    flag = f'  <-- gap {gap:.1f}' if gap >= gap_min else ''
    if flag or '-v' in sys.argv:
        print(f'{s:9.1f} +{d:6.1f} s{streams[st]} {name}{flag}')
print('end', max(last_end.values()), 'streams', streams)
busy = {}
for e in evs:
    st = getattr(e, 'device_resource_id', 0)
    busy[st] = busy.get(st, 0) + (e.time_range.end - e.time_range.start)
print('busy per stream', busy)
# ---- device-level idle analysis (graph replay spreads branches over internal streams, so per-stream gaps mean little)
iv = sorted((e.time_range.start - t0, e.time_range.end - t0, re.sub(r'\(.*', '', re.sub(r'\(anonymous namespace\)::', '', e.name))[:40]) for e in evs)
cur_end, idle, gaps = 0.0, 0.0, []
hist = {}
for s, e_, name in iv:
    if s > cur_end:
        g = s - cur_end
        idle += g
        gaps.append((g, s, name))
        hist[name] = hist.get(name, 0) + g
    cur_end = max(cur_end, e_)
print(f'span {cur_end:.1f} us, device idle (no kernel running) {idle:.1f} us in {len(gaps)} gaps')
print('idle time attributed to the kernel that ENDS the gap:')
for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:14]: print(f'  {v:8.1f} us  {k}')
print('largest gaps:')
for g, s, name in sorted(gaps, reverse=True)[:12]: print(f'  {g:6.1f} us at {s:8.1f} before {name}')
