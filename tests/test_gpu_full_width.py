"""GPU parity at the widths of the configuration the project is about: full 3DiM X-UNet (ch=256, ch_mult=(1,2,2,4),
emb_ch=1024, num_res_blocks=3, 8 heads; BASELINE.json configs[2]/[3]; reference model/xunet.py:205-280 with those attributes).

* operator level: tcgen05 conv / dgrad / wgrad at Cin in {768, 1024, 1536, 2048}, Co in {512, 1024, 2048}, the FiLM 1x1
  GEMM 1024 -> 2048, the fused q|k|v projection at C = 1024, attention at (L=1024, hd=64) and (L=256, hd=128) forward and
  backward -- against fp64 torch restatements evaluated on the same device (independent library kernels, not ours);
* model level: the whole full-3DiM network, B=1, 64x64 (0.88 TFLOP forward): eps_hat and EVERY parameter gradient against
  the CPU oracle, (a) the exact fp32 oracle at the historical bf16-sized tolerances and (b) relative to the bf16 noise floor
  that the rounding-aware oracle (`Bf16Emulation`, rounds where the engine rounds) measures, bias / GroupNorm leaves included.
"""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import xunet_ref as R
from tests.util import rel_l2, to_ref_cfg, np_batch
import novel_view_synthesis_3d_b200 as P

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _conv_ref_gpu(x, w, b, ks):
    """x (N,H,W,Ci) fp64 cuda, w (taps,Ci,Co), stride 1 SAME -> (N,H,W,Co)."""
    if ks == 1:
        return x @ w[0] + b
    Ci, Co = w.shape[1], w.shape[2]
    k = w.reshape(3, 3, Ci, Co).permute(3, 2, 0, 1)
    return F.conv2d(x.permute(0, 3, 1, 2), k, b, padding=1).permute(0, 2, 3, 1)


FULL_WIDTH_CONV = [
    # N, H, Ci, Co, ks, nseg        where it occurs in full-3DiM @128 (per-GPU batch 1)
    (2, 16, 2048, 1024, 3, 1),    # level-3 up block, Conv_0 on concat(1024, 1024)
    (2, 16, 1536, 1024, 3, 1),    # level-3 up block, concat(1024, 512)
    (2, 32, 1536, 512, 3, 1),     # level-2 up block, concat(1024, 512)
    (2, 32, 1024, 512, 3, 1),     # level-2 up block, concat(512, 512)
    (2, 64, 768, 256, 3, 1),      # level-0 up block at reduced side, concat(512, 256)
    (2, 16, 1024, 1024, 3, 1),    # level-3 Conv_1
    (2, 16, 1024, 2048, 1, 1),    # FiLM Dense emb 1024 -> 2 x 1024
    (2, 32, 1024, 1024, 1, 1),    # FiLM Dense emb 1024 -> 2 x 512
    (2, 16, 2048, 1024, 1, 1),    # skip Dense on the concat
    (2, 16, 1024, 3072, 1, 3),    # fused q|k|v projection, C = 1024 (hd 128)
    (2, 32, 512, 1536, 1, 3),     # fused q|k|v projection, C = 512 (hd 64)
    # pair units (two 8 x 16 tiles per pipeline step, conv_tc.cu `m2`): forward and data gradient
    (2, 128, 256, 256, 3, 1),     # level-0 Conv_0 / Conv_1 at 128 x 128
    (4, 64, 512, 512, 3, 1),      # level-1 Conv_0 / Conv_1
    (8, 64, 512, 256, 3, 1),      # Ci != Co
]


@pytest.mark.parametrize('case', FULL_WIDTH_CONV)
def test_conv_tcgen05_full_width(lib, case):
    N, H, Ci, Co, ks, nseg = case
    W = H
    taps, segw = ks * ks, Co // nseg
    g = torch.Generator(device='cuda').manual_seed(hash(case) & 0xFFFF)
    rn = lambda *s: torch.randn(*s, generator=g, device='cuda', dtype=torch.float32)
    x = rn(N, H, W, Ci).to(torch.bfloat16)
    w = rn(nseg, taps, Ci, segw) / math.sqrt(taps * Ci)
    b = rn(Co) * 0.1
    res = rn(N, H, W, Co).to(torch.bfloat16)
    dy = rn(N, H, W, Co).to(torch.bfloat16)
    alpha = 0.7071
    wq = torch.cat([w[s].to(torch.bfloat16).double() for s in range(nseg)], dim=-1)     # (taps, Ci, Co): what the MMA multiplies
    xr = x.double().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    br = b.double().requires_grad_(True)
    out_ref = (_conv_ref_gpu(xr, wr, br, ks) + res.double()) * alpha
    out_ref.backward(dy.double())

    y = torch.zeros(N, H, W, Co, dtype=torch.bfloat16, device='cuda')
    assert lib.xunet_op_conv(1, 1, x.data_ptr(), w.data_ptr(), b.data_ptr(), res.data_ptr(), y.data_ptr(), N, H, W, Ci, Co, ks, 1,
                             nseg, alpha, _stream()) == 0, lib.xunet_last_error()
    torch.cuda.synchronize()
    e_fwd = rel_l2(y.float(), out_ref.detach())
    dx = torch.zeros(N, H, W, Ci, dtype=torch.bfloat16, device='cuda')
    assert lib.xunet_op_conv_dgrad(1, 1, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), N, H, W, Ci, Co, ks, 1, nseg, alpha, 0,
                                   _stream()) == 0, lib.xunet_last_error()
    torch.cuda.synchronize()
    e_dx = rel_l2(dx.float(), xr.grad)
    dw = torch.zeros(nseg, taps, Ci, segw, dtype=torch.float32, device='cuda')
    db = torch.zeros(Co, dtype=torch.float32, device='cuda')
    assert lib.xunet_op_conv_wgrad(1, 1, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, H, W, Ci, Co, ks, 1, nseg,
                                   alpha, _stream()) == 0, lib.xunet_last_error()
    torch.cuda.synchronize()
    dw_ref = torch.stack([wr.grad[..., s * segw:(s + 1) * segw] for s in range(nseg)])
    e_dw, e_db = rel_l2(dw, dw_ref), rel_l2(db, br.grad)
    print(f'full-width conv {case}: fwd {e_fwd:.2e} dgrad {e_dx:.2e} wgrad {e_dw:.2e} dbias {e_db:.2e}')
    assert e_fwd < 6e-3 and e_dx < 6e-3          # one bf16 rounding of the output
    assert e_dw < 1e-4 and e_db < 1e-4           # exact bf16 products, fp32 accumulation


FULL_WIDTH_ATTN = [(2, 1024, 512, 8, 0), (2, 1024, 512, 8, 1), (2, 256, 1024, 8, 0), (2, 256, 1024, 8, 1), (4, 64, 1024, 8, 1)]


@pytest.mark.parametrize('case', FULL_WIDTH_ATTN)
def test_attention_tcgen05_full_width(lib, case):
    """L=1024 / hd=64 (level 2 at 128 px), L=256 / hd=128 (level 3 at 128 px), L=64 / hd=128 (level 3 at 64 px)."""
    N, L, Cc, heads, cross = case
    hd = Cc // heads
    if not bool(L % 128 == 0):
        pytest.skip('L % 128 != 0 runs on the SIMT kernels (covered by test_gpu_ops.py)')
    g = torch.Generator(device='cuda').manual_seed(hash(case) & 0xFFFF)
    rn = lambda *s: torch.randn(*s, generator=g, device='cuda', dtype=torch.float32)
    qkv = (rn(N, L, 3 * Cc) * 1.2).to(torch.bfloat16)
    res = rn(N, L, Cc).to(torch.bfloat16)
    dout = rn(N, L, Cc).to(torch.bfloat16)
    qr = qkv.double().requires_grad_(True)
    q, k, v = [t.reshape(N, L, heads, hd).transpose(1, 2) for t in torch.split(qr, Cc, dim=-1)]     # (N, heads, L, hd)
    if cross:
        perm = torch.arange(N, device='cuda') ^ 1
        k, v = k[perm], v[perm]
    sc = (q / math.sqrt(hd)) @ k.transpose(-1, -2)
    o = (torch.softmax(sc, dim=-1) @ v).transpose(1, 2).reshape(N, L, Cc)
    out_ref = (o + res.double()) / math.sqrt(2)
    out_ref.backward(dout.double())
    od = torch.zeros(N, L, Cc, dtype=torch.bfloat16, device='cuda')
    lse = torch.zeros(N, heads, L, dtype=torch.float32, device='cuda')
    assert lib.xunet_op_attention(1, 1, qkv.data_ptr(), res.data_ptr(), od.data_ptr(), lse.data_ptr(), N, L, Cc, heads, cross,
                                  _stream()) == 0, lib.xunet_last_error()
    torch.cuda.synchronize()
    e_o, e_lse = rel_l2(od.float(), out_ref.detach()), rel_l2(lse, torch.logsumexp(sc, dim=-1).detach())
    scratch = torch.zeros(N * L * (heads + Cc), dtype=torch.float32, device='cuda')
    dqkv = torch.zeros(N, L, 3 * Cc, dtype=torch.bfloat16, device='cuda')
    assert lib.xunet_op_attention_bwd(1, 1, qkv.data_ptr(), res.data_ptr(), od.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                      scratch.data_ptr(), dqkv.data_ptr(), N, L, Cc, heads, cross, _stream()) == 0, lib.xunet_last_error()
    torch.cuda.synchronize()
    gq, gk, gv = [rel_l2(a.float(), b) for a, b in zip(torch.split(dqkv, Cc, dim=-1), torch.split(qr.grad, Cc, dim=-1))]
    print(f'full-width attention {case}: out {e_o:.2e} lse {e_lse:.2e} dq {gq:.2e} dk {gk:.2e} dv {gv:.2e}')
    assert e_lse < 1e-4 and e_o < 8e-3
    assert max(gq, gk, gv) < 4e-2, (gq, gk, gv)


# ------------------------------------------------------------------------------------------------------------------
# whole network at full-3DiM widths
# ------------------------------------------------------------------------------------------------------------------
FULL = dict(ch=256, ch_mult=(1, 2, 2, 4), emb_ch=1024, num_res_blocks=3, attn_resolutions=(8, 16, 32), attn_heads=8, dropout=0.0)


@pytest.fixture(scope='module')
def full64():
    """Full 3DiM at 64x64, B=1: engine run (bf16 product mode) + both oracle variants, computed once for the module."""
    S, B = 64, 1
    torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))     # torch-CPU stops scaling (and regresses) far below 128 threads
    model = P.XUNet(**FULL, dtype='bf16')
    rcfg = to_ref_cfg(model.config)
    # fp32 oracle: 450 M parameters x (value + gradient) and the autograd tape fit comfortably in host memory in fp32
    ref_params = R.init_params(rcfg, S, seed=7, zero_init=False, bias_std=0.1, dtype=torch.float32)
    batch, noise = R.synthetic_batch(B, S, seed=1234, dtype=torch.float32)
    cond = np.ones(B, dtype=np.float32)
    state = P.create_train_state(0, 1, 1e-4, B, S, model=model)
    state.params.flat.copy_(model.flat_from_tree(ref_params, S, B))
    nb = np_batch(batch)
    eps = model.apply({'params': state.params}, nb, cond_mask=cond, train=False).cpu()
    loss, grads = P.apply_model(state, nb['x'], nb['z'], nb['logsnr'], nb['R1'], nb['t1'], nb['R2'], nb['t2'], nb['K'],
                                noise.numpy(), cond_mask=cond)
    torch.cuda.synchronize()
    gflat = {k: v.float().cpu() for k, v in R.flatten(grads).items()}
    out = dict(eps=eps, loss=float(loss), grads=gflat, n_params=int(state.params.flat.numel()))
    for name, emu in (('exact', None), ('emu', R.Bf16Emulation())):
        l, g, e = R.loss_and_grads(ref_params, batch, noise, torch.ones(B), rcfg, train=False, emu=emu)
        out[name] = dict(loss=float(l), grads=g, eps=e)
    return out


def _grad_report(gflat, grads_ref):
    total_ref = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads_ref.values()))
    floor = 1e-3 * total_ref / math.sqrt(len(grads_ref))
    rels = {}
    for k, gr in grads_ref.items():
        err = float(torch.linalg.norm((gflat[k].double() - gr.double()).reshape(-1)))
        rels[k] = err / (float(torch.linalg.norm(gr.double().reshape(-1))) + floor)
    all_g = torch.cat([gflat[k].double().reshape(-1) for k in grads_ref])
    all_r = torch.cat([g.double().reshape(-1) for g in grads_ref.values()])
    return rels, rel_l2(all_g, all_r)


def test_full_3dim_widths_match_exact_oracle(full64):
    """Same bar as the narrow configurations of test_gpu_model.py (exact oracle, bf16-sized tolerances)."""
    r = full64
    assert r['n_params'] == 449877251                               # SURVEY 8(c) item 4, S=64
    e = rel_l2(r['eps'], r['exact']['eps'])
    rels, glob = _grad_report(r['grads'], r['exact']['grads'])
    worst = max(rels.items(), key=lambda kv: kv[1])
    print(f'full-3DiM 64px vs exact fp32 oracle: eps rel-L2 {e:.3e}, loss {r["loss"]:.4f} vs {r["exact"]["loss"]:.4f}, '
          f'grad global {glob:.3e}, worst leaf {worst}')
    assert e < 4e-2
    assert abs(r['loss'] - r['exact']['loss']) / r['exact']['loss'] < 2e-2
    assert glob < 1e-1
    bad = {k: v for k, v in rels.items() if k.endswith('kernel') and v > 2e-1}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]


def test_full_3dim_widths_sit_on_the_bf16_noise_floor(full64):
    """VERDICT r1 item 7 at the real widths: the engine's distance to the exact oracle equals the rounding-aware oracle's own
    (eps_hat, loss, global gradient, every leaf -- bias and GroupNorm leaves included); see
    tests/test_gpu_round2.py::noise_floor_report for why nothing tighter exists for bf16 storage."""
    from tests.test_gpu_round2 import noise_floor_report, worst_leaf_ratio
    r0 = full64
    ex, em = r0['exact'], r0['emu']
    r = noise_floor_report(r0['eps'], r0['loss'], r0['grads'], (ex['loss'], ex['grads'], ex['eps']), (em['loss'], em['grads'], em['eps']))
    worst = sorted(r['leaf_ratio'].items(), key=lambda kv: -kv[1])[:5]
    print(f'full-3DiM 64px bf16 noise floor: eps engine {r["eps_engine"]:.3e} vs floor {r["eps_floor"]:.3e} (engine-vs-emu {r["eps_engine_vs_emu"]:.3e}); '
          f'grad global engine {r["glob_engine"]:.3e} vs floor {r["glob_floor"]:.3e}; loss {r["loss_engine"]:.2e} vs {r["loss_floor"]:.2e}; '
          f'leaf floor median {r["leaf_floor_median"]:.3e}; worst leaf ratios {worst}')
    assert r['eps_engine'] < 1.6 * r['eps_floor'] and r['eps_engine_vs_emu'] < 1.6 * r['eps_floor']
    assert r['glob_engine'] < 1.6 * r['glob_floor']
    assert r['loss_engine'] < max(3 * r['loss_floor'], 2e-3)
    assert worst_leaf_ratio(r) < 3.0, worst
