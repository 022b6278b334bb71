"""Regenerates tests/golden/*.npz from the fp64 CPU oracle (oracle/xunet_ref.py).

    python tests/golden/make_golden.py

The reference itself (JAX/Flax) cannot run in this image, so these vectors pin the ORACLE (and everything
checked against it) against accidental change; they are not outputs of the reference program ("parity unpinned",
see oracle/xunet_ref.py header).  Parameters come from oracle.formula_params (no RNG), inputs from
oracle.synthetic_batch (numpy legacy RandomState, stable across platforms)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import xunet_ref as R  # noqa: E402

CASES = {
    # name: (config, S, B)
    'small64_b2': (R.SMALL, 64, 2),
    'tiny16_b2': (R.RefConfig(ch=32, ch_mult=(1, 2), emb_ch=32, num_res_blocks=1, attn_resolutions=(8, 16), attn_heads=2,
                              dropout=0.0, use_pos_emb=True, use_ref_pose_emb=True), 16, 2),
}


def make(name):
    cfg, S, B = CASES[name]
    params = R.formula_params(cfg, S)
    batch, noise = R.synthetic_batch(B, S, seed=1234)
    cond = torch.tensor([1.0, 0.0][:B] if B <= 2 else [1.0] * B, dtype=torch.float64)
    taps = {}
    eps = R.xunet_forward(params, batch, cond, cfg, train=False, taps=taps)
    loss, grads, _ = R.loss_and_grads(params, batch, noise, cond, cfg, train=False)
    out = {'eps': eps.numpy().astype(np.float32), 'loss': np.float64(loss.item()),
           'cond_mask': cond.numpy()}
    for k, v in taps.items():
        out['tap_mean/' + k] = np.float64(v.mean().item())
        out['tap_std/' + k] = np.float64(v.std().item())
    for k, g in grads.items():
        out['grad_norm/' + k] = np.float64(torch.linalg.norm(g.reshape(-1)).item())
    # a few full gradient leaves (small ones)
    for k in ('GroupNorm_0/GroupNorm_0/scale', 'Conv_1/bias', 'ConditioningProcessor_0/Dense_1/bias'):
        out['grad/' + k] = grads[k].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'eps std', float(eps.std()), 'loss', float(loss))


if __name__ == '__main__':
    for n in CASES:
        make(n)
