"""Device-side forward diffusion: the per-item noise / timestep / z / logsnr / cond_mask work that the
reference does on the host inside SceneInstanceDataset.__getitem__ (dataset/data_loader.py:70-74, 88-110) and train.py:64.

    fd = ForwardDiffusion(engine)                 # uploads the cosine-beta tables once
    fd.sample(x0_target, seed)                    # fills engine.inp['z'], ['noise'], ['logsnr'], ['cond_mask'] on the GPU
Only the clean source / target images and the poses then cross PCIe (no z, no noise)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .sampling import cosine_beta_schedule


class ForwardDiffusion:
    def __init__(self, engine, p_uncond: float = 0.1):
        self.eng, self.lib, self.p_uncond = engine, _lib.load(), float(p_uncond)
        betas = cosine_beta_schedule(1000)
        ac = np.cumprod(1. - betas, axis=0)
        dev = engine.device
        self.sqrt_ac = torch.as_tensor(np.sqrt(ac), dtype=torch.float32).to(dev)               # data_loader.py:73
        self.sqrt_1mac = torch.as_tensor(np.sqrt(1. - ac), dtype=torch.float32).to(dev)        # data_loader.py:74
        self.t = torch.zeros(engine.B, dtype=torch.int32, device=dev)
        self.x0 = torch.zeros_like(engine.inp['z'])

    def sample(self, x0, seed: int, *, t=None, noise=None) -> None:
        """x0: clean target images (B,S,S,3), host or device.  t / noise: optional fixed timesteps / noise (parity tests)."""
        e = self.eng
        src = x0 if isinstance(x0, torch.Tensor) else torch.as_tensor(np.asarray(x0), dtype=torch.float32)
        self.x0.copy_(src.reshape(self.x0.shape).to(torch.float32), non_blocking=True)
        t_ptr = n_ptr = None
        if t is not None:
            self.t.copy_(torch.as_tensor(np.asarray(t), dtype=torch.int32))
            t_ptr = self.t.data_ptr()
        if noise is not None:
            e.inp['noise'].copy_(torch.as_tensor(np.asarray(noise), dtype=torch.float32).reshape(e.inp['noise'].shape))
            n_ptr = e.inp['noise'].data_ptr()
        per = e.S * e.S * 3
        st = torch.cuda.current_stream(e.device).cuda_stream
        _lib.check(self.lib.xunet_forward_diffusion(self.x0.data_ptr(), n_ptr, t_ptr, int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                    self.sqrt_ac.data_ptr(), self.sqrt_1mac.data_ptr(), self.p_uncond,
                                                    e.inp['z'].data_ptr(), e.inp['noise'].data_ptr(), e.inp['logsnr'].data_ptr(),
                                                    self.t.data_ptr() if t is None else None, e.inp['cond_mask'].data_ptr(),
                                                    e.B, per, st), 'xunet_forward_diffusion')
