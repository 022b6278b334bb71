"""Full 3DiM X-UNet (ch=256, (1,2,2,4), emb 1024, nrb 3, heads 8) at 128x128: does it run, how fast, how much memory."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import novel_view_synthesis_3d_b200 as P
from bench import make_host_batches
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 128
model = P.XUNet.from_config(P.FULL_3DIM)
t0 = time.time()
state = P.create_train_state(0, 1, 1e-4, B, S, model=model, zero_init=False)
print('init s', time.time() - t0, 'params', state.params.flat.numel())
eng = model.engine(B, S, True)
print('workspace GB', eng.ws_bytes / 1e9, 'kernels fwd/bwd', eng.count_kernels(state.params.flat))
host = make_host_batches(1, B, S, 1)
eng.load_inputs(host[0][0], cond_mask=np.ones(B, np.float32), noise=host[0][1])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(3):
    ev[0].record(); eng.forward(state.params.flat, train=True, seed=1); ev[1].record(); eng.backward(state.params.flat); ev[2].record()
    torch.cuda.synchronize()
    print(f'it{it}: fwd {ev[0].elapsed_time(ev[1]):.2f} ms  bwd {ev[1].elapsed_time(ev[2]):.2f} ms  loss {float(eng.loss):.4f} eps_std {float(eng.eps.std()):.4f} finite {bool(torch.isfinite(eng.grads).all())}')
fl = 3155.6e9 * B
print('fwd TFLOP/s', fl / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e12, ' train TFLOP/s', 3 * fl / (ev[0].elapsed_time(ev[2]) * 1e-3) / 1e12)
print('max mem GB', torch.cuda.max_memory_allocated() / 1e9)
