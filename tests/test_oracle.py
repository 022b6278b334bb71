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


# ---------------------------------------------------------------------------------------------------------------
# Independent cross-checks of the oracle's primitives ("stop the oracle pinning itself").  Each compares
# oracle/xunet_ref.py against a DIFFERENT implementation of the same library semantics: torch.nn.functional's own
# group_norm / scaled_dot_product_attention / avg_pool2d / interpolate, and a shift-and-add numpy convolution that
# re-derives XLA's SAME padding from its definition instead of calling F.pad + F.conv2d.
# ---------------------------------------------------------------------------------------------------------------
import torch.nn.functional as F


@pytest.mark.parametrize('C,H', [(32, 4), (96, 6), (256, 8)])
def test_primitive_group_norm_vs_torch_functional(C, H):
    g = torch.Generator().manual_seed(C)
    h = torch.randn(3, 2, H, H + 2, C, generator=g, dtype=torch.float64) * 2.0 + 0.7
    p = {'GroupNorm_0': {'scale': torch.randn(C, generator=g, dtype=torch.float64), 'bias': torch.randn(C, generator=g, dtype=torch.float64)}}
    ours = R.group_norm(h, p)
    # joint-frame statistics == torch's GroupNorm on the (B, C, F*H, W) view (all non-batch, non-channel axes pooled)
    x = h.permute(0, 4, 1, 2, 3).reshape(3, C, 2 * H, H + 2)
    ref = F.group_norm(x, 32, p['GroupNorm_0']['scale'], p['GroupNorm_0']['bias'], eps=1e-6)
    ref = ref.reshape(3, C, 2, H, H + 2).permute(0, 2, 3, 4, 1)
    assert rel_l2(ours, ref) < 1e-12


@pytest.mark.parametrize('heads,hd,Lq,Lk', [(4, 16, 64, 64), (8, 64, 16, 32), (1, 32, 10, 7)])
def test_primitive_attention_vs_torch_sdpa(heads, hd, Lq, Lk):
    g = torch.Generator().manual_seed(heads * 100 + hd)
    C = heads * hd
    pl = {f'DenseGeneral_{i}': {'kernel': torch.randn(C, heads, hd, generator=g, dtype=torch.float64) / math.sqrt(C),
                                'bias': torch.randn(heads, hd, generator=g, dtype=torch.float64) * 0.1} for i in range(3)}
    q_in = torch.randn(2, Lq, C, generator=g, dtype=torch.float64)
    kv_in = torch.randn(2, Lk, C, generator=g, dtype=torch.float64)
    ours = R.attn_layer(q_in, kv_in, pl, heads)                                   # (B, Lq, heads, hd)

    def proj(x, i):   # DenseGeneral == a plain matmul with the (C, heads*hd) reshaped kernel
        w, b = pl[f'DenseGeneral_{i}']['kernel'].reshape(C, C), pl[f'DenseGeneral_{i}']['bias'].reshape(C)
        return (x @ w + b).reshape(x.shape[0], x.shape[1], heads, hd).transpose(1, 2)   # (B, heads, L, hd)
    ref = F.scaled_dot_product_attention(proj(q_in, 0), proj(kv_in, 1), proj(kv_in, 2))  # scale 1/sqrt(hd), no mask
    assert rel_l2(ours, ref.transpose(1, 2)) < 1e-12


def test_primitive_resamplers_vs_torch_functional():
    g = torch.Generator().manual_seed(4)
    h = torch.randn(2, 2, 6, 8, 5, generator=g, dtype=torch.float64)
    x = h.reshape(4, 6, 8, 5).permute(0, 3, 1, 2)
    up = F.interpolate(x, scale_factor=2, mode='nearest').permute(0, 2, 3, 1).reshape(2, 2, 12, 16, 5)
    assert torch.equal(R.nearest_neighbor_upsample(h), up)
    dn = F.avg_pool2d(x, 2).permute(0, 2, 3, 1).reshape(2, 2, 3, 4, 5)
    assert rel_l2(R.avgpool_downsample(h), dn) < 1e-15
    odd = torch.randn(1, 2, 7, 5, 3, generator=g, dtype=torch.float64)              # VALID pooling drops the odd row / column
    dn = F.avg_pool2d(odd.reshape(2, 7, 5, 3).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).reshape(1, 2, 3, 2, 3)
    assert rel_l2(R.avgpool_downsample(odd), dn) < 1e-15


def _conv_same_numpy(x, k, b, s):
    """XLA 'SAME' convolution from its definition: out[o] = sum_t x[o*s + t - lo] k[t], out = ceil(in/s),
    total pad = max((out-1)*s + 3 - in, 0), lo = total // 2; out-of-range taps contribute zero."""
    N, H, W, Ci = x.shape
    Co = k.shape[-1]
    Ho, Wo = -(-H // s), -(-W // s)
    lo_h = max((Ho - 1) * s + 3 - H, 0) // 2
    lo_w = max((Wo - 1) * s + 3 - W, 0) // 2
    y = np.zeros((N, Ho, Wo, Co))
    for dy in range(3):
        for dx in range(3):
            for oy in range(Ho):
                iy = oy * s + dy - lo_h
                if iy < 0 or iy >= H:
                    continue
                for ox in range(Wo):
                    ix = ox * s + dx - lo_w
                    if ix < 0 or ix >= W:
                        continue
                    y[:, oy, ox, :] += x[:, iy, ix, :] @ k[dy, dx]
    return y + b


@pytest.mark.parametrize('H,W,s', [(8, 8, 1), (8, 8, 2), (16, 16, 4), (16, 16, 8), (7, 9, 1), (7, 9, 2), (10, 6, 4)])
def test_primitive_conv_same_padding_vs_definition(H, W, s):
    g = torch.Generator().manual_seed(H * 10 + s)
    h = torch.randn(1, 2, H, W, 3, generator=g, dtype=torch.float64)
    k = torch.randn(1, 3, 3, 3, 4, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    ours = R.conv_1x3x3(h, k, b, stride=s)
    ref = _conv_same_numpy(h.reshape(2, H, W, 3).numpy(), k[0].numpy(), b.numpy(), s)
    assert ours.shape[2:4] == ref.shape[1:3]
    assert rel_l2(ours.reshape(ref.shape), ref) < 1e-13


def test_primitive_swish_film_dense_vs_torch_functional():
    g = torch.Generator().manual_seed(6)
    x = torch.randn(50, dtype=torch.float64, generator=g)
    assert rel_l2(R.swish(x), F.silu(x)) < 1e-15
    h = torch.randn(1, 2, 3, 3, 8, dtype=torch.float64, generator=g)
    emb = torch.randn(1, 2, 3, 3, 6, dtype=torch.float64, generator=g)
    p = {'Dense_0': {'kernel': torch.randn(6, 16, dtype=torch.float64, generator=g), 'bias': torch.randn(16, dtype=torch.float64, generator=g)}}
    e = F.linear(F.silu(emb), p['Dense_0']['kernel'].t(), p['Dense_0']['bias'])
    assert rel_l2(R.film(h, emb, p), h * (1 + e[..., :8]) + e[..., 8:]) < 1e-14          # first half = scale (jnp.split)


def test_zero_init_rule_excludes_the_pose_embedding_conv():
    """'ConditioningProcessor_0/Conv_1/kernel' is the level-1 pose conv (model/xunet.py:197-202, default
    lecun_normal) and must not be caught by the out_init_scale filter (model/xunet.py:85-89,276-280)."""
    from novel_view_synthesis_3d_b200.xunet import is_zero_init_kernel
    shapes = R.param_shapes(R.FULL, 64)
    zero = [k for k in shapes if k.endswith('kernel') and R._is_zero_init(k)]
    assert zero and all(k == 'Conv_1/kernel' or k.split('/')[-3].startswith('ResnetBlock_') for k in zero)
    assert [k for k in shapes if k.endswith('kernel') and is_zero_init_kernel(k)] == zero      # product rule == oracle rule
    p = R.init_params(TINY, 16, seed=0, zero_init=True, flat=True)
    for i in range(len(TINY.ch_mult)):
        assert float(p[f'ConditioningProcessor_0/Conv_{i}/kernel'].abs().max()) > 0
    n_res = sum(1 for k in p if k.endswith('/Conv_1/kernel') and 'ResnetBlock_' in k)
    assert sum(1 for k, v in p.items() if k.endswith('kernel') and float(v.abs().max()) == 0) == n_res + 1


def test_bf16_emulation_is_a_small_structured_perturbation():
    """Bf16Emulation rounds where the engine rounds: the result stays within bf16-sized distance of the exact oracle,
    gradients flow through every leaf, and the exact path is untouched (no emulation object = identity)."""
    p = R.init_params(TINY, 16, seed=7, zero_init=False, bias_std=0.1)
    batch, noise = R.synthetic_batch(2, 16, seed=5)
    one = torch.ones(2)
    exact = R.xunet_forward(p, batch, one, TINY)
    emu = R.xunet_forward(p, batch, one, TINY, emu=R.Bf16Emulation())
    d = rel_l2(emu, exact)
    assert 1e-4 < d < 5e-2, d
    # every value the engine would hand back is bf16-representable
    assert torch.equal(emu, emu.to(torch.bfloat16).to(emu.dtype))
    l0, g0, _ = R.loss_and_grads(p, batch, noise, one, TINY, train=False)
    l1, g1, _ = R.loss_and_grads(p, batch, noise, one, TINY, train=False, emu=R.Bf16Emulation())
    assert abs(float(l0) - float(l1)) / float(l0) < 2e-2
    a0 = torch.cat([v.reshape(-1) for v in g0.values()])
    a1 = torch.cat([v.reshape(-1) for v in g1.values()])
    assert 1e-4 < rel_l2(a1, a0) < 1e-1
    assert all(float(v.abs().max()) > 0 for v in g1.values())
