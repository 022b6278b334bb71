"""Shared helpers for the parity tests (tests may import the oracle; the product may not)."""
import numpy as np
import torch

from oracle import xunet_ref as R


def to_ref_cfg(cfg) -> R.RefConfig:
    return R.RefConfig(ch=cfg.ch, ch_mult=tuple(cfg.ch_mult), emb_ch=cfg.emb_ch, num_res_blocks=cfg.num_res_blocks,
                       attn_resolutions=tuple(cfg.attn_resolutions), attn_heads=cfg.attn_heads, dropout=cfg.dropout,
                       use_pos_emb=cfg.use_pos_emb, use_ref_pose_emb=cfg.use_ref_pose_emb)


def rel_l2(a, b) -> float:
    a = torch.as_tensor(np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a), dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b), dtype=torch.float64).reshape(-1)
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


def np_batch(batch):
    return {k: v.numpy() for k, v in batch.items()}


# numpy replica of xu_keep() in csrc/common.cuh
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def keep_mask(seed: int, op_index: int, shape, rate: float) -> np.ndarray:
    """xu_keep(): one 64-bit hash per 4 consecutive elements, 16-bit uniforms."""
    n = int(np.prod(shape))
    n4 = (n + 3) // 4
    with np.errstate(over='ignore'):
        idx4 = np.arange(n4, dtype=np.uint64)
        z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(op_index + 1) * np.uint64(0xD1B54A32D192ED03) + idx4
        z = _mix64(z)
    thr = np.uint64(int(np.float32(rate) * np.float32(65536.0)))
    lanes = np.stack([((z >> np.uint64(16 * j)) & np.uint64(0xFFFF)) >= thr for j in range(4)], axis=1).reshape(-1)
    return lanes[:n].reshape(shape)
