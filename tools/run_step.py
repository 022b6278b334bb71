"""A few eager (no CUDA graph) training steps of the bench workload -- target for ncu -k regex:<kernel>."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import novel_view_synthesis_3d_b200 as P
from bench import make_host_batches
import os
B, S = int(os.environ.get('XU_B', 8)), int(os.environ.get('XU_S', 64))      # XU_MODEL=full XU_B=2 XU_S=128: the full 3DiM model
full = os.environ.get('XU_MODEL', 'small') == 'full'
model = P.XUNet.from_config(P.XUNetConfig(**{**P.FULL_3DIM.__dict__, 'dtype': 'bf16'})) if full else P.XUNet(dtype='bf16')
state = P.create_train_state(0, 1, 1e-4, B, S, model=model, init_on_device=full)
step = P.TrainStep(state, use_graph=False)
host = make_host_batches(1, B, S, 1234)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    step(host[0][0], host[0][1], cond_mask=np.ones(B, np.float32))
torch.cuda.synchronize()
print('done')
