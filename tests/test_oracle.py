"""Pins for the CPU oracle: the known-answer / metamorphic properties derivable from the reference source
 and the committed fp64 golden vectors.  CPU only."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import xunet_ref as R
from tests.util import rel_l2

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TINY = R.RefConfig(ch=32, ch_mult=(1, 2), emb_ch=32, num_res_blocks=1, attn_resolutions=(8, 16), attn_heads=2,
                   dropout=0.0)


def test_param_counts():  # item 4
    assert R.param_count(R.SMALL, 64) == 1054211
    assert R.param_count(R.SMALL, 128) == 902915
    assert R.param_count(R.FULL, 128) == 438831363
    assert R.param_count(R.FULL, 64) == 449877251


def test_zero_init_gives_zero_eps():  # item 1 (model/xunet.py:85-89, 276-280)
    p = R.init_params(TINY, 16, seed=3, zero_init=True)
    batch, _ = R.synthetic_batch(2, 16, seed=5)
    out = R.xunet_forward(p, batch, torch.ones(2), TINY)
    assert float(out.abs().max()) == 0.0


def test_cond_mask_zero_is_pose_invariant():  # item 2 (model/xunet.py:178-179)
    p = R.formula_params(TINY, 16)
    b1, _ = R.synthetic_batch(2, 16, seed=5)
    b2 = dict(b1)
    o, _ = R.synthetic_batch(2, 16, seed=99)
    for k in ('R1', 't1', 'R2', 't2'):
        b2[k] = o[k]
    z = torch.zeros(2)
    assert rel_l2(R.xunet_forward(p, b1, z, TINY), R.xunet_forward(p, b2, z, TINY)) < 1e-12
    one = torch.ones(2)
    assert rel_l2(R.xunet_forward(p, b1, one, TINY), R.xunet_forward(p, b2, one, TINY)) > 1e-4


def test_frame_swap_equivariance():  # item 3
    p = R.formula_params(TINY, 16)
    b, _ = R.synthetic_batch(2, 16, seed=7)
    sw = dict(b, x=b['z'], z=b['x'], R1=b['R2'], t1=b['t2'], R2=b['R1'], t2=b['t1'])
    one = torch.ones(2)
    a = R.xunet_forward(p, b, one, TINY, both_frames=True)
    c = R.xunet_forward(p, sw, one, TINY, both_frames=True)
    assert rel_l2(a[:, 1], c[:, 0]) < 1e-10 and rel_l2(a[:, 0], c[:, 1]) < 1e-10


def test_posenc_ddpm_known_answer():  # item 5 / SURVEY a3
    t = torch.tensor([2 * math.atan(math.exp(0.0)) / math.pi], dtype=torch.float64)   # logsnr 0 -> 0.5
    e = R.posenc_ddpm(t, 32, 1.)[0]
    assert np.allclose(e[:3].numpy(), [-0.46777181, 0.39658617, 0.9399987], atol=1e-7)
    assert np.allclose(e[16:19].numpy(), [-0.88384927, 0.9179975, -0.34117803], atol=1e-7)


def test_posenc_nerf_layout():  # item 5 / SURVEY a4: [x, sin(xb) scale-major xyz-minor, sin(xb+pi/2)]
    x = torch.tensor([[0.1, -0.2, 0.3]], dtype=torch.float64)
    e = R.posenc_nerf(x, 0, 15)[0]
    assert e.shape == (93,)
    x32 = x.float()
    assert torch.allclose(e[:3], x32[0].double())
    assert abs(float(e[3 + 3 * 2 + 1]) - math.sin(float(x32[0, 1]) * 4)) < 1e-12      # scale 2^2, component y
    # phase channel: the argument xb + pi/2 is an fp32 add in the reference (ulp 2^-9 at ~4915)
    arg = float(x32[0, 2] * 2 ** 14 + torch.tensor(math.pi / 2, dtype=torch.float32))
    assert abs(float(e[3 + 45 + 3 * 14 + 2]) - math.sin(arg)) < 1e-12
    assert abs(arg - (0.3 * 2 ** 14 + math.pi / 2)) < 2e-3
    assert R.posenc_nerf(x, 0, 8).shape[-1] == 51 and 93 + 51 == R.POSE_EMB_DIM


def test_same_padding_geometry():  # item 6 / SURVEY a2
    assert R.same_pad(64, 3, 1) == (1, 1)
    assert R.same_pad(64, 3, 2) == (0, 1)
    assert R.same_pad(64, 3, 4) == (0, 0)
    assert R.same_pad(64, 3, 8) == (0, 0)
    h = torch.arange(2 * 8 * 8, dtype=torch.float64).reshape(1, 2, 8, 8, 1)
    k = torch.zeros(1, 3, 3, 1, 1, dtype=torch.float64)
    k[0, 0, 0, 0, 0] = 1.0   # picks the top-left tap: with stride 2 pad (0,1) that is pixel (2i, 2j)
    y = R.conv_1x3x3(h, k, torch.zeros(1, dtype=torch.float64), stride=2)
    assert y.shape == (1, 2, 4, 4, 1) and float(y[0, 0, 1, 1, 0]) == float(h[0, 0, 2, 2, 0])


def test_groupnorm_is_joint_over_frames():  # item 7 / F6
    p = {'GroupNorm_0': {'scale': torch.ones(32, dtype=torch.float64), 'bias': torch.zeros(32, dtype=torch.float64)}}
    g = torch.Generator().manual_seed(0)
    h = torch.randn(1, 2, 4, 4, 32, generator=g, dtype=torch.float64)
    h2 = h.clone()
    h2[:, 0] += 3.0
    a, b = R.group_norm(h, p), R.group_norm(h2, p)
    assert float((a[:, 1] - b[:, 1]).abs().max()) > 0.1


def test_attention_properties():  # item 8
    g = torch.Generator().manual_seed(1)
    C, heads = 32, 2
    pl = {f'DenseGeneral_{i}': {'kernel': torch.randn(C, heads, C // heads, generator=g, dtype=torch.float64) * 0.2,
                                'bias': torch.randn(heads, C // heads, generator=g, dtype=torch.float64) * 0.1} for i in range(3)}
    q = torch.randn(1, 16, C, generator=g, dtype=torch.float64)
    kv = torch.randn(1, 16, C, generator=g, dtype=torch.float64)
    perm = torch.randperm(16, generator=g)
    assert rel_l2(R.attn_layer(q, kv[:, perm], pl, heads), R.attn_layer(q, kv, pl, heads)) < 1e-12
    p = {'GroupNorm_0': {'GroupNorm_0': {'scale': torch.ones(C, dtype=torch.float64), 'bias': torch.zeros(C, dtype=torch.float64)}},
         'AttnLayer_0': pl}
    h = torch.randn(1, 1, 4, 4, C, generator=g, dtype=torch.float64).expand(1, 2, 4, 4, C).contiguous()
    assert rel_l2(R.attn_block(h, p, 'cross', heads), R.attn_block(h, p, 'self', heads)) < 1e-12


def test_loss_and_adam():  # item 9 / F4
    g = torch.Generator().manual_seed(2)
    e = torch.randn(2, 4, 4, 3, generator=g, dtype=torch.float64)
    n = torch.randn(2, 4, 4, 3, generator=g, dtype=torch.float64)
    assert abs(float(R.loss_fn(e, n)) - math.sqrt(float(((e - n) ** 2).sum()))) < 1e-12
    p = torch.randn(10, dtype=torch.float64, generator=g)
    gr = torch.randn(10, dtype=torch.float64, generator=g)
    p2, m, v = R.adam_update(p, gr, torch.zeros(10, dtype=torch.float64), torch.zeros(10, dtype=torch.float64), 1, lr=1e-4)
    assert torch.allclose(p2 - p, -1e-4 * torch.sign(gr), atol=1e-9)


def test_schedule_constants():  # item 10
    t = R.schedule_tables()
    assert t['betas'][0] == pytest.approx(4.128422482196914e-05, rel=1e-12)
    assert t['betas'][-1] == 0.9999
    assert t['alphas_cumprod'][0] == pytest.approx(0.999958715775178, rel=1e-14)
    assert t['alphas_cumprod'][-1] == pytest.approx(2.4288e-10, rel=1e-4)
    assert R.logsnr_schedule_cosine(0.0) == pytest.approx(20.0)
    assert abs(R.logsnr_schedule_cosine(0.5)) < 1e-12
    assert R.logsnr_schedule_cosine(0.999) == pytest.approx(-12.8555, abs=1e-4)


def test_resblock_at_init_is_scaled_skip():  # item 1b: zero Conv_1 => block = skip(h_in)/sqrt2
    cfg = TINY
    p = R.init_params(cfg, 16, seed=1, zero_init=True)['XUNetBlock_0']['ResnetBlock_0']
    g = torch.Generator().manual_seed(3)
    h = torch.randn(1, 2, 8, 8, 32, generator=g, dtype=torch.float64)
    emb = torch.randn(1, 2, 8, 8, 32, generator=g, dtype=torch.float64)
    out = R.resnet_block(h, emb, p, features=32)
    assert rel_l2(out, h / math.sqrt(2)) < 1e-12


@pytest.mark.parametrize('name', ['tiny16_b2', 'small64_b2'])
def test_oracle_matches_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLD, 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cfg, S, B = mg.CASES[name]
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    params = R.formula_params(cfg, S)
    batch, noise = R.synthetic_batch(B, S, seed=1234)
    cond = torch.from_numpy(gold['cond_mask'])
    taps = {}
    eps = R.xunet_forward(params, batch, cond, cfg, train=False, taps=taps)
    assert rel_l2(eps, gold['eps']) < 1e-6
    for k, v in taps.items():
        assert float(v.mean()) == pytest.approx(float(gold['tap_mean/' + k]), rel=1e-8, abs=1e-10)
    # fp32 oracle (the timed CPU baseline) agrees with fp64 to well under the 1e-3 budget
    p32 = R.formula_params(cfg, S, dtype=torch.float32)
    b32, _ = R.synthetic_batch(B, S, seed=1234, dtype=torch.float32)
    assert rel_l2(R.xunet_forward(p32, b32, cond.float(), cfg), gold["eps"]) < 1e-3
    if name == 'tiny16_b2':
        loss, grads, _ = R.loss_and_grads(params, batch, noise, cond, cfg, train=False)
        assert float(loss) == pytest.approx(float(gold['loss']), rel=1e-9)
        for k, g in grads.items():
            assert float(torch.linalg.norm(g.reshape(-1))) == pytest.approx(float(gold['grad_norm/' + k]), rel=1e-6, abs=1e-12)
