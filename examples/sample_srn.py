"""sampling.py look-alike (reference: sampling.py:55-167): restore a Flax-format checkpoint, draw one batch, run the CFG
ancestral sampler, write PNGs (the reference blocks in cv2.imshow, which cannot run headless).

    python examples/sample_srn.py cars_train_val --ckpt checkpoints --steps 256 --out results/
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import novel_view_synthesis_3d_b200 as P
from novel_view_synthesis_3d_b200.srn_data import SRNScenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('folder')
    ap.add_argument('--ckpt', default='checkpoints')
    ap.add_argument('--steps', type=int, default=1000)       # the reference runs all 1000 steps, sampling.py:128
    ap.add_argument('--w', type=float, default=3.0)          # sampling.py:133
    ap.add_argument('--views', type=int, default=1)
    ap.add_argument('--side', type=int, default=64)
    ap.add_argument('--out', default='results')
    ap.add_argument('--orbit', type=int, default=0,
                    help='3DiM stochastic conditioning: generate this many further views autoregressively (the poses of the '
                         'next pairs drawn from the loader), each denoising step conditioned on a random view of the pool')
    a = ap.parse_args()
    params = P.checkpoint.restore_checkpoint(a.ckpt, prefix='model')
    if params is None:
        raise FileNotFoundError('Checkpoint does not exist')                                          # sampling.py:111-112
    ds = SRNScenes(a.folder, img_sidelength=a.side, max_observations_per_instance=50)
    batch = next(ds.batches(a.views))
    model = P.XUNet()
    sampler = P.Sampler(model, params, a.views, a.side, steps=a.steps, w=a.w)
    os.makedirs(a.out, exist_ok=True)
    if a.orbit > 0:
        it = ds.batches(a.views)
        more = [next(it) for _ in range(a.orbit)]
        tp = {'R': np.stack([b['R2'] for b in more], 1), 't': np.stack([b['t2'] for b in more], 1)}
        seq = sampler.sample_views(batch['x'][:, None], {'R': batch['R1'][:, None], 't': batch['t1'][:, None]}, batch['K'], tp)
        for i, views in enumerate(seq.cpu().numpy()):
            for j, img in enumerate(views):
                P.sampling.save_view(os.path.join(a.out, f'object_{i}_view_{j}.png'), img)
        return
    z = sampler.sample(batch)
    for i, img in enumerate(z.cpu().numpy()):
        P.sampling.save_view(os.path.join(a.out, f'view_{i}.png'), img)                               # z/2 + 0.5, sampling.py:153
        P.sampling.save_view(os.path.join(a.out, f'source_{i}.png'), batch['x'][i])


if __name__ == '__main__':
    main()
