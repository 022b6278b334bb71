"""train.py look-alike on the B200 path (reference: train.py:78-176).  Differences from the reference script: imports,
a sharded batch under torchrun, device-side forward diffusion, and a Flax-format checkpoint every `save_every` steps.

    python examples/train_srn.py cars_train_val --batch 8 --side 64 --steps 1000
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_srn.py cars_train_val
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import novel_view_synthesis_3d_b200 as P
from novel_view_synthesis_3d_b200 import dist as xdist
from novel_view_synthesis_3d_b200.srn_data import SRNScenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('folder')
    ap.add_argument('--batch', type=int, default=2)          # Trainer(train_batch_size=2), train.py:84
    ap.add_argument('--side', type=int, default=64)          # img_sidelength=64, :88
    ap.add_argument('--lr', type=float, default=1e-4)        # train_lr, :85
    ap.add_argument('--steps', type=int, default=100000)     # train_num_steps, :86
    ap.add_argument('--save-every', type=int, default=1000)  # :87
    ap.add_argument('--dtype', default='bf16')
    a = ap.parse_args()
    local = xdist.init_from_env('nccl')
    rank, world = xdist.rank(), xdist.world_size()
    ds = SRNScenes(a.folder, img_sidelength=a.side, max_observations_per_instance=50, seed=rank)      # train.py:99-104
    model = P.XUNet(dtype=a.dtype)
    state = P.create_train_state(0, 1, a.lr, a.batch, a.side, model=model)
    step = P.TrainStep(state)
    fd = P.ForwardDiffusion(step.eng)
    stream = ds.batches(a.batch)
    for it in range(a.steps):
        b = next(stream)
        fd.sample(b['target'], seed=it * world + rank)        # z, noise, logsnr, cond_mask on the GPU
        dev = step.eng.inp
        batch = {k: b[k] for k in ('x', 'R1', 't1', 'R2', 't2', 'K')}
        batch.update(z=dev['z'], logsnr=dev['logsnr'])
        loss = step(batch, dev['noise'], cond_mask=dev['cond_mask'])
        if rank == 0:
            if it % 50 == 0:
                print(f'{it}: {float(loss):.4f}')                                                      # train.py:157
            if it % a.save_every == 0:
                P.checkpoint.save_checkpoint('checkpoints/', state.params, step=it, prefix='model', add_device_axis=True)
    if rank == 0:
        print('training completed')


if __name__ == '__main__':
    main()
