"""CPU tests of the host-side logic: schedule tables vs the oracle's restatement of sampling.py, seeds, the numpy
replica of the dropout hash, and the data-parallel helpers under a 2-process gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import xunet_ref as R
from novel_view_synthesis_3d_b200 import Schedule, cosine_beta_schedule, logsnr_schedule_cosine
from novel_view_synthesis_3d_b200 import dist as xdist
from novel_view_synthesis_3d_b200.xunet import _seed_of, _nest, _flatten
from tests.util import keep_mask


def test_schedule_matches_oracle_tables():
    t, s = R.schedule_tables(), Schedule(1000)
    assert np.array_equal(cosine_beta_schedule(1000), t['betas'])
    for k in ('sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod', 'posterior_mean_coef1', 'posterior_mean_coef2',
              'posterior_log_variance_clipped'):
        assert np.allclose(getattr(s, k), t[k], rtol=1e-13, atol=0), k
    assert logsnr_schedule_cosine(0.25) == R.logsnr_schedule_cosine(0.25)


def test_respaced_schedule():
    s = Schedule(256)
    assert len(s) == 256 and s.timesteps[0] == 0 and s.timesteps[-1] == 999
    full = np.cumprod(1 - cosine_beta_schedule(1000))
    assert np.allclose(s.alphas_cumprod, full[s.timesteps])
    assert np.all(s.posterior_variance >= 0) and s.posterior_variance[0] == 0


def test_seed_and_tree_helpers():
    assert _seed_of(5) == 5 and _seed_of(None, 3) == 3
    assert _seed_of(np.array([0, 7], dtype=np.uint32)) == 7
    tree = {'a': {'b': 1, 'c': {'d': 2}}, 'e': 3}
    assert _nest(_flatten(tree)) == tree


def test_dropout_hash_statistics():
    m = keep_mask(123, 4, (64, 64, 32), 0.1)
    assert abs(m.mean() - 0.9) < 5e-3
    assert not np.array_equal(m, keep_mask(124, 4, (64, 64, 32), 0.1))
    assert not np.array_equal(m, keep_mask(123, 5, (64, 64, 32), 0.1))
    assert np.array_equal(m, keep_mask(123, 4, (64, 64, 32), 0.1))


def test_shard_batch():
    batch = {'x': np.arange(8 * 3).reshape(8, 3), 'logsnr': np.arange(8)}
    parts = [xdist.shard_batch(batch, r, 4) for r in range(4)]
    assert all(p['x'].shape == (2, 3) for p in parts)
    assert np.array_equal(np.concatenate([p['logsnr'] for p in parts]), np.arange(8))
    with pytest.raises(ValueError):
        xdist.shard_batch(batch, 0, 3)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        flat = torch.full((1000,), float(rank + 1))
        xdist.broadcast_params(flat)                       # rank 0's parameters everywhere
        assert torch.all(flat == 1.0)
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        xdist.allreduce_sum_(g, bucket_elems=300)          # bucketed all-reduce of the flat gradient buffer
        assert torch.allclose(g, torch.arange(1000, dtype=torch.float32) * sum(range(1, world + 1)))
        t = xdist.max_over_ranks(float(rank + 10))
        assert t == world + 9
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2_data_parallel_helpers():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert dict(out) == {0: 1, 1: 1}
