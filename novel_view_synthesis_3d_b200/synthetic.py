"""Synthetic SRN-shaped batches: the input contract of dataset/data_loader.py:102-113 without the dataset.

x, x0 ~ U(-1,1); noise ~ N(0,1); t ~ U{0..999}; z = sqrt(abar_t) x0 + sqrt(1-abar_t) noise; logsnr = cosine(t/1000);
camera centres uniform on the r=1.3 sphere looking at the origin (cam->world R, t); K = [[f,0,S/2],[0,f,S/2],[0,0,1]],
f = 131.25*S/128 (SRN cars intrinsics rescaled as dataset/util.py:64-67).  dtypes as the loader hands them over: x float32,
z / noise float64."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .sampling import cosine_beta_schedule, logsnr_schedule_cosine


def _look_at(c: np.ndarray) -> np.ndarray:
    fwd = -c / np.linalg.norm(c)
    right = np.cross(fwd, np.array([0., 0., 1.]))
    if np.linalg.norm(right) < 1e-6:
        right = np.array([1., 0., 0.])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    return np.stack([right, down, fwd], axis=1)          # columns = camera x (right), y (down), z (forward) in world


def synthetic_batch(B: int, S: int, seed: int = 1234) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
    rng = np.random.RandomState(seed)
    ac = np.cumprod(1. - cosine_beta_schedule(1000), axis=0)
    x = rng.uniform(-1, 1, (B, S, S, 3)).astype(np.float32)
    x0 = rng.uniform(-1, 1, (B, S, S, 3))
    noise = rng.randn(B, S, S, 3)
    t = rng.randint(0, 1000, (B,))
    z = np.sqrt(ac[t])[:, None, None, None] * x0 + np.sqrt(1. - ac[t])[:, None, None, None] * noise
    Rs, ts = [], []
    for _ in range(2):
        c = rng.randn(B, 3)
        c = 1.3 * c / np.linalg.norm(c, axis=1, keepdims=True)
        Rs.append(np.stack([_look_at(ci) for ci in c]))
        ts.append(c)
    f = 131.25 * S / 128.
    K = np.tile(np.array([[f, 0, S / 2.], [0, f, S / 2.], [0, 0, 1.]])[None], (B, 1, 1))
    batch = dict(x=x, z=z, logsnr=logsnr_schedule_cosine(t / 1000.0), R1=Rs[0], t1=ts[0], R2=Rs[1], t2=ts[1], K=K)
    return batch, noise
