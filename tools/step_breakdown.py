"""Where does one training step spend its time?  (host staging, graph launch, device time)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import novel_view_synthesis_3d_b200 as P
from bench import make_host_batches

B, S = 8, 64
model = P.XUNet(dtype='bf16')
state = P.create_train_state(0, 1, 1e-4, B, S, model=model)
step = P.TrainStep(state, use_graph=True)
host = make_host_batches(2, B, S, 1234)
mask = np.ones(B, np.float32)
for i in range(5):
    step(host[i % 2][0], host[i % 2][1], cond_mask=mask)
torch.cuda.synchronize()
eng = step.eng

def t_host(fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3

print('load_inputs  host/total ms', t_host(lambda: eng.load_inputs(host[0][0], cond_mask=mask, noise=host[0][1])))
print('graph_fb.replay host/total ms', t_host(lambda: step.graph_fb.replay()))
print('graph_opt.replay host/total ms', t_host(lambda: step.graph_opt.replay()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step.graph_fb.replay()
e1.record(); torch.cuda.synchronize()
print('graph_fb device ms', e0.elapsed_time(e1) / 20)
def full():
    l = step(host[0][0], host[0][1], cond_mask=mask); return float(l)
print('full step with loss.item() host/total ms', t_host(full))
def nosync():
    step(host[0][0], host[0][1], cond_mask=mask)
print('full step no sync host/total ms', t_host(nosync))
# forward / backward split without graphs
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.synchronize()
e[0].record(); eng.forward(state.params.flat, train=True); e[1].record(); eng.backward(state.params.flat); e[2].record(); torch.cuda.synchronize()
print('no-graph forward ms', e[0].elapsed_time(e[1]), 'backward ms', e[1].elapsed_time(e[2]))
