"""Time the 1x1 (FiLM dense) data-gradient at the full-model shape, with and without accumulation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from novel_view_synthesis_3d_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
bf = torch.bfloat16
N, H, Ci, Co = 8, 128, 1024, 512      # forward conv Ci -> Co; dgrad: dy (Co) -> dx (Ci)
dy = torch.randn(N, H, H, Co, device='cuda').to(bf)
dx = torch.zeros(N, H, H, Ci, device='cuda', dtype=bf)
w = torch.randn(Ci * Co, device='cuda') * 0.02
for acc in (0, 1):
    f = lambda: lib.xunet_op_conv_dgrad(1, 1, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), N, H, H, Ci, Co, 1, 1, 1, 1.0, acc, st)
    for _ in range(3): assert f() == 0, lib.xunet_last_error()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print('accumulate', acc, f'{e0.elapsed_time(e1) / 10 * 1e3:.1f} us')
