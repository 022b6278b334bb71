"""Round-2 GPU tests: the second ray convention and the explicit-rays entry, reference_quirks
(train.py:64,66 frozen mask + dropout key), the pinned staging ring, the data-parallel step (2 ranks: identical
parameters on every rank, step == hand-averaged gradient), rank-decorrelated dropout, and the bf16 product mode against the
rounding-aware oracle on the narrow configurations."""
import dataclasses
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import xunet_ref as R
from tests.util import rel_l2, to_ref_cfg, np_batch, keep_mask
import novel_view_synthesis_3d_b200 as P

pytestmark = pytest.mark.gpu

TINY = dict(ch=32, ch_mult=(1, 2), emb_ch=32, num_res_blocks=1, attn_resolutions=(8, 16), attn_heads=2, dropout=0.0)
SMALL = dict(ch=32, ch_mult=(1, 2), emb_ch=32, num_res_blocks=2, attn_resolutions=(8, 16, 32), attn_heads=4, dropout=0.1)
FOUR = dict(ch=64, ch_mult=(1, 2, 2, 4), emb_ch=128, num_res_blocks=1, attn_resolutions=(8, 16), attn_heads=8, dropout=0.0)


def _setup(cfgd, S, B, dtype, **kw):
    model = P.XUNet(**cfgd, dtype=dtype, **kw)
    rcfg = to_ref_cfg(model.config)
    ref_params = R.formula_params(rcfg, S) if dtype == 'fp32' else R.init_params(rcfg, S, seed=7, zero_init=False, bias_std=0.1)
    tree = model.tree_from_flat(model.flat_from_tree(ref_params, S, B), S, B)
    batch, noise = R.synthetic_batch(B, S, seed=1234)
    return model, rcfg, ref_params, tree, batch, noise


# ---- rays ------------------------------------------------------------------------------------------------------------
def _asym_batch(batch, S):
    """SRN's K is symmetric (fx = fy, cx = cy), which makes the two pixel conventions exact transposes of each other;
    an asymmetric K makes a convention mix-up visible in every pixel."""
    b = dict(batch)
    K = batch['K'].clone()
    K[:, 0, 0] *= 1.25
    K[:, 0, 2] += 0.11 * S
    K[:, 1, 2] -= 0.07 * S
    b['K'] = K
    return b


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_opencv_uv_ray_convention_matches_oracle(dtype):
    S, B = 16, 2
    model, rcfg, ref_params, tree, batch, _ = _setup(TINY, S, B, dtype, ray_convention='opencv_uv')
    batch = _asym_batch(batch, S)
    cond = torch.ones(B, dtype=torch.float64)
    taps = {}
    ref = R.xunet_forward(ref_params, batch, cond, rcfg, ray_convention='opencv_uv', taps=taps)
    other = R.xunet_forward(ref_params, batch, cond, rcfg, ray_convention='v3d130_ij')
    eps = model.apply({'params': tree}, np_batch(batch), cond_mask=cond.numpy(), train=False)
    tol = 1e-3 if dtype == 'fp32' else 4e-2
    eng = model.engine(B, S, False)
    for i in range(2):
        t = taps[f'pose_emb_{i}']
        assert rel_l2(eng.read_tap(f'pose_emb_{i}'), t.reshape(-1, *t.shape[-3:])) < 2 * tol
    assert rel_l2(eps, ref) < tol
    assert rel_l2(other, ref) > 10 * tol          # the two conventions really differ on these inputs
    # and the default convention is the other one
    model_ij, *_ = _setup(TINY, S, B, dtype)
    eps_ij = model_ij.apply({'params': tree}, np_batch(batch), cond_mask=cond.numpy(), train=False)
    assert rel_l2(eps_ij, other) < tol


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_explicit_rays_entry_matches_oracle(dtype):
    """xunet_batch.rays / XUNet.apply(..., rays=): the rays a reference-side v3d.Camera(...).rays() produces are fed in
    directly, so parity does not depend on the library's restatement of visu3d (model/xunet.py:159-161,166-168)."""
    S, B = 16, 2
    model, rcfg, ref_params, tree, batch, _ = _setup(TINY, S, B, dtype)
    batch = _asym_batch(batch, S)
    cond = torch.tensor([1.0, 0.0], dtype=torch.float64)
    tol = 1e-3 if dtype == 'fp32' else 4e-2
    nb = np_batch(batch)
    for conv in ('opencv_uv', 'v3d130_ij'):
        r1 = R.camera_rays(batch['R1'], batch['t1'], batch['K'], S, S, conv)
        r2 = R.camera_rays(batch['R2'], batch['t2'], batch['K'], S, S, conv)
        ref = R.xunet_forward(ref_params, batch, cond, rcfg, rays=(r1, r2))
        rays = torch.stack([torch.cat([r1[0], r1[1]], -1), torch.cat([r2[0], r2[1]], -1)], dim=1)      # (B,2,S,S,6)
        # R, t, K are ignored when rays are given: scramble them to prove it
        scr = dict(nb, R1=nb['R2'], t1=nb['t2'] * 0 + 5.0, K=nb['K'] * 3.0)
        eps = model.apply({'params': tree}, scr, cond_mask=cond.numpy(), train=False, rays=rays.numpy())
        assert rel_l2(eps, ref) < tol, conv
    # generic rays (per-pixel origins, not a pinhole camera): the entry is a plain tensor input
    g = torch.Generator().manual_seed(3)
    pos = torch.randn(B, 2, S, S, 3, generator=g, dtype=torch.float64) * 0.5
    d = torch.randn(B, 2, S, S, 3, generator=g, dtype=torch.float64)
    d = d / torch.linalg.norm(d, dim=-1, keepdim=True)
    ref = R.xunet_forward(ref_params, batch, torch.ones(B), rcfg, rays=((pos[:, 0], d[:, 0]), (pos[:, 1], d[:, 1])))
    eps = model.apply({'params': tree}, nb, cond_mask=np.ones(B), train=False, rays=torch.cat([pos, d], -1).numpy())
    assert rel_l2(eps, ref) < (2e-3 if dtype == 'fp32' else 4e-2)
    # the entry is per call: the next plain call uses R, t, K again
    eps_plain = model.apply({'params': tree}, nb, cond_mask=np.ones(B), train=False)
    assert rel_l2(eps_plain, R.xunet_forward(ref_params, batch, torch.ones(B), rcfg)) < tol
    with pytest.raises(ValueError):
        model.apply({'params': tree}, nb, cond_mask=np.ones(B), train=False, rays=np.zeros((B, 2, S, S, 5)))


# ---- reference_quirks ------------------------------------------------------------------------------------------------
def test_reference_quirks_freeze_cond_mask_and_dropout_key():
    """train.py:64,66: cond_mask and PRNGKey(0) are evaluated once at trace time -> the same mask and the same dropout
    pattern at every step.  The default (fresh per step) differs."""
    S, B = 16, 8
    cfgd = dict(TINY, dropout=0.3)
    batch, noise = R.synthetic_batch(B, S, seed=5)
    nb = np_batch(batch)
    args = (nb['x'], nb['z'], nb['logsnr'], nb['R1'], nb['t1'], nb['R2'], nb['t2'], nb['K'], noise.numpy())

    def run(quirks):
        model = P.XUNet(**cfgd, dtype='fp32')
        st = P.create_train_state(0, 123, 1e-3, B, S, model=model, zero_init=False, reference_quirks=quirks)
        eng = model.engine(B, S, True)
        out = []
        for _ in range(3):
            loss, grads = P.apply_model(st, *args)
            out.append((float(loss), eng.inp['cond_mask'].cpu().numpy().copy(), int(eng.seed.item()), grads.flat.clone()))
            st = dataclasses.replace(st, step=st.step + 1)                 # advance the step WITHOUT touching the parameters
        return out

    q = run(True)
    assert all(np.array_equal(q[0][1], r[1]) for r in q) and all(r[2] == 0 for r in q)             # frozen mask, key 0
    assert all(abs(r[0] - q[0][0]) < 1e-5 * q[0][0] for r in q)                                    # identical function
    assert all(rel_l2(r[3], q[0][3]) < 1e-4 for r in q)
    # the frozen pattern is exactly the hash mask of seed 0
    model = P.XUNet(**cfgd, dtype='fp32')
    rcfg = to_ref_cfg(model.config)
    st = P.create_train_state(0, 123, 1e-3, B, S, model=model, zero_init=False, reference_quirks=True)
    loss, _ = P.apply_model(st, *args)
    mask = model.engine(B, S, True).inp['cond_mask'].cpu().double()
    ref_params = {k: v.double().cpu() for k, v in R.flatten(st.params).items()}
    lref, _, _ = R.loss_and_grads(R.nest(ref_params), batch, noise, mask, rcfg, train=True,
                                  drop_mask_fn=lambda idx, shape: torch.from_numpy(keep_mask(0, idx, shape, 0.3)))
    assert abs(float(loss) - float(lref)) / float(lref) < 2e-4
    d = run(False)
    assert len({r[2] for r in d}) == 3                                                           # fresh dropout seed per step
    assert max(abs(r[0] - d[0][0]) for r in d) > 1e-4 * d[0][0]
    # fused step: the device-side seed stays 0 under quirks and advances otherwise
    for quirks in (True, False):
        model = P.XUNet(**cfgd, dtype='fp32')
        st = P.create_train_state(0, 123, 1e-3, B, S, model=model, zero_init=False, reference_quirks=quirks)
        step = P.TrainStep(st)
        for _ in range(3):
            step(nb, noise.numpy())
        assert int(step.eng.seed.item()) == (0 if quirks else 3)


# ---- pinned staging ring ---------------------------------------------------------------------------------------------
def test_back_to_back_steps_do_not_race_the_pinned_staging_buffer():
    """the host runs many steps ahead of the GPU; every step must train on ITS batch.  A long kernel
    queue is built first so that all H2D copies of the loop below are still pending while the host rewrites its buffers."""
    S, B, n = 16, 4, 24
    batches = [R.synthetic_batch(B, S, seed=100 + i) for i in range(n)]
    cond = np.ones(B, dtype=np.float32)

    def run(sync):
        model = P.XUNet(**TINY, dtype='fp32')
        st = P.create_train_state(0, 1, 0.0, B, S, model=model, zero_init=False)     # lr 0: every step sees the same weights
        step = P.TrainStep(st)
        step(np_batch(batches[0][0]), batches[0][1].numpy(), cond_mask=cond)          # capture
        torch.cuda.synchronize()
        losses = torch.zeros(n, device='cuda')
        if not sync:
            spin = torch.randn(8192, 8192, device='cuda')
            for _ in range(12):
                spin @ spin                                                            # ~0.2 s of queued device work (fp32 GEMMs)
        for i, (b, nz) in enumerate(batches):
            losses[i] = step(np_batch(b), nz.numpy(), cond_mask=cond)
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return losses.cpu().numpy()

    a, b = run(True), run(False)
    assert len(np.unique(np.round(a, 4))) == n                                        # distinct batches -> distinct losses
    assert np.allclose(a, b, rtol=1e-5), np.abs(a - b).max()


# ---- data parallel ---------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, backend, use_graph, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    if backend == 'nccl':
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', dev))
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        S, B = 16, 2
        cfgd = dict(TINY, dropout=0.2)
        model = P.XUNet(**cfgd, dtype='fp32')
        # rank-dependent init on purpose: create_train_state must broadcast rank 0's parameters
        st = P.create_train_state(rank + 1, 7, 1e-3, B, S, model=model, zero_init=False)
        p0 = st.params.flat.clone()
        step = P.TrainStep(st, use_graph=use_graph, bucket_mb=0.05)        # tiny buckets: several all-reduces per step
        masks = []
        for it in range(3):
            batch, noise = R.synthetic_batch(B, S, seed=1000 + 10 * it + rank)        # every rank its own shard
            step(np_batch(batch), noise.numpy())
            if it == 0:
                g_sum = step.eng.grads.clone()           # all-reduced (summed) gradient of step 0
                seed0 = int(step.eng.seed.item())
            masks.append(step.eng.inp['cond_mask'].cpu().numpy().copy())
        torch.cuda.synchronize()
        # local (un-reduced) gradient of step 0 with the same seed / mask, through the two-call API without collectives
        model2 = P.XUNet(**cfgd, dtype='fp32')
        eng2 = model2.engine(B, S, True)
        batch, noise = R.synthetic_batch(B, S, seed=1000 + rank)
        eng2.load_inputs(np_batch(batch), cond_mask=masks[0], noise=noise.numpy())
        eng2.forward(p0, train=True, seed=seed0)
        _, g_local = eng2.backward(p0)
        torch.cuda.synchronize()
        torch.save(dict(params=st.params.flat.cpu(), p0=p0.cpu(), g_sum=g_sum.cpu(), g_local=g_local.clone().cpu(), seed0=seed0,
                        mask0=masks[0], mode=step.mode, buckets=len(step.reducer.ranges)), os.path.join(out_dir, f'r{rank}.pt'))
        # captured graphs hold NCCL kernels of this communicator: release them before it is torn down
        import gc
        torch.cuda.synchronize()
        del step, eng2
        gc.collect()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('backend,use_graph', [('gloo', True), ('gloo', False), ('nccl', True), ('nccl', False)])
def test_data_parallel_step_two_ranks(tmp_path, backend, use_graph):
    """north_star: per-GPU forward/backward + ONE all-reduce of the gradient bucket (bucketed + overlapped here).  After 3
    steps every rank holds bit-identical parameters; step 0's reduced gradient is the sum of the two local gradients; the
    dropout seeds / cond_masks of the ranks differ."""
    import torch.multiprocessing as mp
    if backend == 'nccl' and torch.cuda.device_count() < 2:
        pytest.skip('NCCL needs one GPU per rank (gloo variant covers the host logic on a single GPU)')
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, backend, use_graph, str(tmp_path))) for r in range(world)]
    [p.start() for p in procs]
    [p.join(240) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    r0, r1 = [torch.load(os.path.join(str(tmp_path), f'r{r}.pt'), weights_only=False) for r in range(world)]
    assert torch.equal(r0['p0'], r1['p0'])                                   # rank 0's init everywhere
    assert torch.equal(r0['params'], r1['params'])                           # bit-identical after 3 steps
    assert not torch.equal(r0['params'], r0['p0'])
    assert torch.equal(r0['g_sum'], r1['g_sum'])
    assert rel_l2(r0['g_sum'], r0['g_local'] + r1['g_local']) < 1e-4         # == hand-summed (fp32 atomics: ~1e-5 run to run)
    assert r0['seed0'] != r1['seed0'] and (r0['seed0'] >> 40) == 0 and (r1['seed0'] >> 40) == 1
    expect = {('nccl', True): 'one_graph', ('gloo', True): 'two_graphs_eager_collectives'}.get((backend, use_graph), 'eager')
    assert r0['mode'] == expect, r0['mode']
    if expect != 'two_graphs_eager_collectives':      # hook-driven modes: several buckets were reduced while the backward ran
        assert r0['buckets'] >= 2, r0['buckets']


# ---- bf16 product mode vs the rounding-aware oracle on the narrow configurations --------------------------------------
def noise_floor_report(eng_eps, eng_loss, eng_grads, exact, emu):
    """bf16 storage makes the network chaotic at the 1e-2 level: the SAME rounding-aware oracle evaluated with fp32 instead of
    fp64 accumulation already lands 1.0e-2 away from itself (measured, small model 64 px), so no implementation can be held
    tighter than that against any other.  What CAN be held tightly is the statistics: the engine is 'exact + bf16 storage
    noise' exactly like the rounding-aware oracle, so its distance to the exact oracle must match the oracle's own
    emulation-vs-exact distance -- for eps_hat, the loss, the global gradient and every single leaf (bias and GroupNorm
    leaves included: their larger cancellation noise shows up in the floor too).  Returns the ratios."""
    (l0, g0, e0), (l1, g1, e1) = exact, emu
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in g0.values()))
    floor_abs = 1e-3 * total / math.sqrt(len(g0))
    leaf = lambda a, b: float(torch.linalg.norm((a.double().cpu() - b.double()).reshape(-1))) / (float(torch.linalg.norm(b.double().reshape(-1))) + floor_abs)
    fl = {k: leaf(g1[k], g0[k]) for k in g0}
    en = {k: leaf(eng_grads[k], g0[k]) for k in g0}
    med = float(np.median(list(fl.values())))
    cat = lambda d: torch.cat([d[k].double().cpu().reshape(-1) for k in g0])
    rep = dict(eps_floor=rel_l2(e1, e0), eps_engine=rel_l2(eng_eps, e0), eps_engine_vs_emu=rel_l2(eng_eps, e1),
               glob_floor=rel_l2(cat(g1), cat(g0)), glob_engine=rel_l2(cat(eng_grads), cat(g0)),
               loss_floor=abs(float(l1) - float(l0)) / float(l0), loss_engine=abs(eng_loss - float(l0)) / float(l0),
               leaf_ratio={k: en[k] / max(fl[k], med) for k in g0}, leaf_floor_median=med)
    return rep


def worst_leaf_ratio(r):
    """Largest engine-noise / oracle-noise ratio over the gradient leaves.  The key-projection bias (AttnLayer DenseGeneral_1/bias) is
    left out of the 3x bound and held to 6x: its true gradient is EXACTLY zero (a constant added to every key shifts all scores of a
    query equally and softmax ignores it), so both numbers are pure cancellation noise of sum_keys dK and their ratio depends on the
    summation order alone (measured 2.1-3.1 across kernels plans on the same inputs)."""
    zero_grad = lambda k: k.endswith('DenseGeneral_1/bias')
    a = max(v for k, v in r['leaf_ratio'].items() if not zero_grad(k))
    b = max([v for k, v in r['leaf_ratio'].items() if zero_grad(k)] or [0.0])
    return max(a, b / 2.0)


@pytest.mark.parametrize('cfgd,S,B', [(SMALL, 64, 2), (FOUR, 64, 1)])
def test_bf16_mode_sits_on_the_bf16_noise_floor(cfgd, S, B):
    """VERDICT r1 item 7 (what 'parity' means for the product dtype): see noise_floor_report."""
    cfgd = dict(cfgd, dropout=0.0)
    model, rcfg, ref_params, tree, batch, noise = _setup(cfgd, S, B, 'bf16')
    cond = np.array(([1.0, 0.0] * B)[:B])
    state = P.create_train_state(0, 1, 1e-4, B, S, model=model)
    state.params.flat.copy_(tree.flat)
    nb = np_batch(batch)
    exact = R.loss_and_grads(ref_params, batch, noise, torch.from_numpy(cond), rcfg, train=False)
    emu = R.loss_and_grads(ref_params, batch, noise, torch.from_numpy(cond), rcfg, train=False, emu=R.Bf16Emulation())
    eps = model.apply({'params': state.params}, nb, cond_mask=cond, train=False)
    loss, grads = P.apply_model(state, nb['x'], nb['z'], nb['logsnr'], nb['R1'], nb['t1'], nb['R2'], nb['t2'], nb['K'],
                                noise.numpy(), cond_mask=cond)
    r = noise_floor_report(eps, float(loss), R.flatten(grads), exact, emu)
    worst = sorted(r['leaf_ratio'].items(), key=lambda kv: -kv[1])[:4]
    print(f'bf16 noise floor [{cfgd["ch"]}ch S={S}]: eps engine {r["eps_engine"]:.3e} vs floor {r["eps_floor"]:.3e} (engine-vs-emu {r["eps_engine_vs_emu"]:.3e}); '
          f'grad global engine {r["glob_engine"]:.3e} vs floor {r["glob_floor"]:.3e}; loss {r["loss_engine"]:.2e} vs {r["loss_floor"]:.2e}; '
          f'worst leaf ratios {worst}')
    assert r['eps_engine'] < 1.6 * r['eps_floor'] and r['eps_engine_vs_emu'] < 1.6 * r['eps_floor']
    assert r['glob_engine'] < 1.6 * r['glob_floor']
    assert r['loss_engine'] < max(3 * r['loss_floor'], 2e-3)
    assert worst_leaf_ratio(r) < 3.0, worst


# ---- A/B of the fused GroupNorm paths (plans are chosen per xunet_create from the environment) --------------------------
def _grads_with_env(env, cfgd, S, B):
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        model, rcfg, ref_params, tree, batch, noise = _setup(cfgd, S, B, 'bf16')
        state = P.create_train_state(0, 1, 1e-4, B, S, model=model)
        state.params.flat.copy_(tree.flat)
        nb = np_batch(batch)
        cond = np.ones(B)
        eps = model.apply({'params': state.params}, nb, cond_mask=cond, train=False).clone()
        loss, grads = P.apply_model(state, nb['x'], nb['z'], nb['logsnr'], nb['R1'], nb['t1'], nb['R2'], nb['t2'], nb['K'],
                                    noise.numpy(), cond_mask=cond)
        eng = model.engine(B, S, True)
        nf, nbk = eng.count_kernels(state.params.flat)
        return eps, float(loss), grads.flat.clone(), (nf, nbk)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize('cfgd,S,B', [(dict(SMALL, dropout=0.0), 64, 2), (FOUR, 64, 1)])
def test_fused_groupnorm_paths_agree_with_the_separate_kernels(cfgd, S, B):
    """Producer-emitted statistics (conv / attention epilogues) and the data-gradient epilogue that does the first pass of the
    GroupNorm backward replace stand-alone kernels.  bf16 storage is chaotic at the 1e-2 level (noise_floor_report), so the
    three plans are not compared with each other but each against the exact oracle: all of them must sit on the same bf16
    noise floor, with strictly fewer launches."""
    base = _grads_with_env({'XUNET_GN_SEPARATE_STATS': '1', 'XUNET_GN_BWD_FUSED': '0'}, cfgd, S, B)
    fwd = _grads_with_env({'XUNET_GN_SEPARATE_STATS': None, 'XUNET_GN_BWD_FUSED': '0'}, cfgd, S, B)
    both = _grads_with_env({'XUNET_GN_SEPARATE_STATS': None, 'XUNET_GN_BWD_FUSED': '1'}, cfgd, S, B)
    print('kernels (fwd, bwd): separate', base[3], 'fused stats', fwd[3], 'fused stats + backward', both[3])
    assert fwd[3][0] < base[3][0] and both[3][1] < fwd[3][1]
    model, rcfg, ref_params, tree, batch, noise = _setup(cfgd, S, B, 'bf16')
    cond = torch.ones(B, dtype=torch.float64)
    exact = R.loss_and_grads(ref_params, batch, noise, cond, rcfg, train=False)
    emu = R.loss_and_grads(ref_params, batch, noise, cond, rcfg, train=False, emu=R.Bf16Emulation())
    names = list(exact[1].keys())
    for tag, (eps, loss, gflat_dev, _) in (('separate', base), ('fused stats', fwd), ('fused stats + backward', both)):
        spec = model.param_spec(S, B)
        g = {k: gflat_dev[spec[k][1]: spec[k][1] + int(np.prod(spec[k][0]))].reshape(spec[k][0]) for k in names}
        r = noise_floor_report(eps, loss, g, exact, emu)
        worst = sorted(r['leaf_ratio'].items(), key=lambda kv: -kv[1])[:3]
        print(f'  {tag}: eps {r["eps_engine"]:.3e} (floor {r["eps_floor"]:.3e}), grad global {r["glob_engine"]:.3e} (floor {r["glob_floor"]:.3e}), '
              f'worst leaf ratios {worst}')
        assert r['eps_engine'] < 1.6 * r['eps_floor'] and r['glob_engine'] < 1.6 * r['glob_floor'], tag
        assert worst_leaf_ratio(r) < 3.0, (tag, worst)


WIDE = dict(ch=256, ch_mult=(1,), emb_ch=256, num_res_blocks=2, attn_resolutions=(), attn_heads=8, dropout=0.0)


def test_pair_unit_convolutions_agree_with_the_single_tile_pipeline():
    """conv_tc.cu `m2`: at 256 channels and 128 x 128 the 3x3 convolutions (forward with fused statistics, data gradients with the
    fused GroupNorm / FiLM backward epilogues) run two 8 x 16 tiles per pipeline step.  XUNET_CONV_M2=0 restores the single-tile
    pipeline that the oracle tests pin; the two must agree to the run-to-run noise of bf16 storage (atomics order in the fused
    statistics), far below the O(1) error of a wrong tile mapping."""
    S, B = 128, 1
    a = _grads_with_env({'XUNET_CONV_M2': '0'}, WIDE, S, B)
    a2 = _grads_with_env({'XUNET_CONV_M2': '0'}, WIDE, S, B)
    b = _grads_with_env({'XUNET_CONV_M2': '1'}, WIDE, S, B)
    noise_eps = rel_l2(a2[0].float(), a[0].float().cpu())
    noise_g = rel_l2(a2[2], a[2].cpu())
    e_eps = rel_l2(b[0].float(), a[0].float().cpu())
    e_g = rel_l2(b[2], a[2].cpu())
    print(f'pair units vs single tiles: eps {e_eps:.3e} (run-to-run {noise_eps:.3e}), grads {e_g:.3e} (run-to-run {noise_g:.3e}), '
          f'loss {b[1]:.6f} vs {a[1]:.6f}')
    assert e_eps < max(3 * noise_eps, 2e-2) and e_g < max(3 * noise_g, 3e-2)
    assert abs(b[1] - a[1]) < 2e-2 * abs(a[1])
