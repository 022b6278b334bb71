"""Host-side mirror of the reference's Flax API for the X-UNet (model/xunet.py:205-280).

    model = XUNet()                                            # model/xunet.py:205-215 attributes
    params = model.init({'params': key, 'dropout': key}, sample, cond_mask=np.zeros(B), train=True)['params']
    eps = model.apply({'params': params}, batch, cond_mask=mask, train=True, rngs={'dropout': key})

All compute runs in libxunet_b200.so (hand-written sm_100a CUDA); PyTorch is used only for device memory,
streams and torch.distributed.  There is no CPU / eager fallback: without the library or a GPU this raises.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

POSE_EMB_DIM = 144
BATCH_KEYS = ('x', 'z', 'logsnr', 'R1', 't1', 'R2', 't2', 'K')
INPUT_ORDER = BATCH_KEYS + ('cond_mask', 'noise')     # segment order of Engine.inp_all


@dataclass(frozen=True)
class XUNetConfig:
    """The nine XUNet attributes of model/xunet.py:207-215 (+ the two B200-side knobs)."""
    ch: int = 32
    ch_mult: Tuple[int, ...] = (1, 2)
    emb_ch: int = 32
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (8, 16, 32)
    attn_heads: int = 4
    dropout: float = 0.1
    use_pos_emb: bool = False
    use_ref_pose_emb: bool = False
    dtype: str = 'bf16'                 # 'bf16' (activations bf16, fp32 accumulate) | 'fp32' (exact-fp32 verify mode)
    ray_convention: str = 'v3d130_ij'   # see SURVEY 8(c): visu3d 1.3.0 pixel convention

    def c_struct(self) -> _lib.XunetConfig:
        c = _lib.XunetConfig()
        c.ch = self.ch
        c.n_levels = len(self.ch_mult)
        for i, m in enumerate(self.ch_mult):
            c.ch_mult[i] = m
        c.emb_ch = self.emb_ch
        c.num_res_blocks = self.num_res_blocks
        c.n_attn_resolutions = len(self.attn_resolutions)
        for i, r in enumerate(self.attn_resolutions):
            c.attn_resolutions[i] = r
        c.attn_heads = self.attn_heads
        c.dropout = float(self.dropout)
        c.use_pos_emb = int(self.use_pos_emb)
        c.use_ref_pose_emb = int(self.use_ref_pose_emb)
        c.ray_convention = _lib.RAYS[self.ray_convention]
        return c


# named presets for BASELINE.json's configs
SMALL = XUNetConfig()
FULL_3DIM = XUNetConfig(ch=256, ch_mult=(1, 2, 2, 4), emb_ch=1024, num_res_blocks=3, attn_resolutions=(8, 16, 32),
                        attn_heads=8)


class ParamTree(dict):
    """Nested dict of parameter leaves in the Flax tree shape.  `.flat` is the single fp32
    device buffer all leaves are views of (one gradient bucket / one Adam launch)."""
    flat: torch.Tensor = None
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], int]]" = None


def _nest(flat: Dict[str, torch.Tensor]) -> dict:
    tree: dict = {}
    for k, v in flat.items():
        node = tree
        parts = k.split('/')
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = v
    return tree


def _flatten(tree: dict, prefix='') -> Dict[str, object]:
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + k + '/'))
        else:
            out[prefix + k] = v
    return out


def is_zero_init_kernel(name: str) -> bool:
    """out_init_scale() leaves (model/xunet.py:11-12): every ResnetBlock's Conv_1 (:85-89) and the top-level Conv_1
    (:276-280).  NOT 'ConditioningProcessor_0/Conv_1/kernel' -- the level-1 pose-embedding conv (:197-202) keeps the
    default lecun_normal although its auto-name also ends in Conv_1."""
    return name.endswith('Conv_1/kernel') and not name.startswith('ConditioningProcessor_0/')


def _seed_of(key, default=0) -> int:
    if key is None:
        return default
    if isinstance(key, (int, np.integer)):
        return int(key)
    arr = np.asarray(key).reshape(-1)           # e.g. a jax PRNGKey-like uint32[2]
    hi = int(arr[-2]) if arr.size > 1 else 0
    return ((hi << 32) | (int(arr[-1]) & 0xFFFFFFFF)) & 0x7FFFFFFFFFFFFFFF


class Engine:
    """One compiled execution plan: (config, B, S, training).  Owns the workspace and static I/O buffers."""
    PIN_SLOTS = int(__import__('os').environ.get('XUNET_PIN_SLOTS', '3'))

    def __init__(self, cfg: XUNetConfig, B: int, S: int, training: bool, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('novel_view_synthesis_3d_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.cfg, self.B, self.S, self.training = cfg, B, S, training
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.dtype_code = {'fp32': _lib.DTYPE_F32, 'bf16': _lib.DTYPE_BF16}[cfg.dtype]
        self.act_dtype = torch.float32 if cfg.dtype == 'fp32' else torch.bfloat16
        h = C.c_void_p()
        cs = cfg.c_struct()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.xunet_create(C.byref(cs), B, S, self.dtype_code, int(training), C.byref(h)), 'xunet_create')
        self.h = h
        self.nparams = int(self.lib.xunet_param_count(h))
        self.spec = OrderedDict()
        name, ndim, shape, off = C.c_char_p(), C.c_int(), (C.c_longlong * 5)(), C.c_longlong()
        for i in range(self.lib.xunet_param_leaves(h)):
            _lib.check(self.lib.xunet_param_leaf(h, i, C.byref(name), C.byref(ndim), C.byref(shape), C.byref(off)), 'param_leaf')
            self.spec[name.value.decode()] = (tuple(int(shape[k]) for k in range(ndim.value)), int(off.value))
        self.ws_bytes = int(self.lib.xunet_workspace_bytes(h))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        f32 = dict(dtype=torch.float32, device=self.device)
        # all step inputs live in ONE device buffer (and one pinned mirror), 256-byte aligned segments in INPUT_ORDER: a
        # training step stages its batch with a single H2D copy instead of ten small ones (each costs ~7 us of stream time)
        shapes = {'x': (B, S, S, 3), 'z': (B, S, S, 3), 'logsnr': (B,), 'R1': (B, 3, 3), 't1': (B, 3), 'R2': (B, 3, 3), 't2': (B, 3),
                  'K': (B, 3, 3), 'cond_mask': (B,), 'noise': (B, S, S, 3)}
        self._seg, off = {}, 0
        for k in INPUT_ORDER:
            n = int(np.prod(shapes[k]))
            self._seg[k] = (off, n)
            off += (n + 63) // 64 * 64
        self.inp_all = torch.zeros(off, **f32)
        self.inp = {k: self.inp_all[o:o + n].view(shapes[k]) for k, (o, n) in self._seg.items()}
        self.inp['cond_mask'].fill_(1.0)
        # pinned staging ring: the host may run several steps ahead of the GPU (a step is a few ms of device time, the host
        # side far less), so a slot is rewritten only after the H2D copy that last read it has completed (one event per slot)
        self._pin_ring = [torch.zeros(off, dtype=torch.float32).pin_memory() for _ in range(self.PIN_SLOTS)]
        self._pin_np = [t.numpy() for t in self._pin_ring]      # numpy views of the same pinned memory
        self._pin_events = [None] * self.PIN_SLOTS
        self._pin_next = 0
        self.rays = None          # optional (B,2,S,S,6) device tensor: explicit-rays entry (set_rays)
        self.eps = torch.zeros(B, S, S, 3, **f32)
        self.loss = torch.zeros(1, **f32)
        self.seed = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.grads = torch.zeros(self.nparams, **f32) if training else None
        self.cbatch = _lib.XunetBatch(**{k: self.inp[k].data_ptr() for k in BATCH_KEYS + ('cond_mask',)}, rays=None)
        self._taps = None
        self._bucket_cb = None    # keeps the ctypes callback object alive while it is installed

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.xunet_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- host -> device staging (pinned memory, async on the current stream) -------------------------
    def load_inputs(self, batch: dict, cond_mask=None, noise=None) -> int:
        """Copies one batch dict (numpy / torch, any float dtype) into the static device buffers.
        Returns the number of host->device bytes moved."""
        nbytes = 0
        items = [(k, batch[k]) for k in BATCH_KEYS]
        if cond_mask is not None:
            items.append(('cond_mask', cond_mask))
        if noise is not None:
            items.append(('noise', noise))
        staged = []
        slot, pin = None, None
        for k, v in items:
            dst = self.inp[k]
            if isinstance(v, torch.Tensor) and v.is_cuda:
                dst.copy_(v.reshape(dst.shape).to(torch.float32), non_blocking=True)
                continue
            src = (v.detach().float() if v.dtype == torch.bfloat16 else v.detach()).numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            if tuple(src.shape) != tuple(dst.shape) and src.size != dst.numel():
                raise ValueError(f"batch['{k}'] has shape {tuple(src.shape)}, expected {tuple(dst.shape)}")
            if pin is None:
                slot = self._pin_next
                self._pin_next = (slot + 1) % self.PIN_SLOTS
                if self._pin_events[slot] is not None:
                    self._pin_events[slot].synchronize()      # the copy that last read this slot has finished
                pin = self._pin_ring[slot]
            o, n = self._seg[k]
            # float64 -> float32 down-cast (as JAX does with x64 off) straight into pinned memory.  numpy, not torch: a torch CPU
            # copy_ of ~1e5 elements fans out over every host core (3.5 ms on 8 threads against 0.1 ms single-threaded)
            np.copyto(self._pin_np[slot][o:o + n], src.reshape(-1), casting='unsafe')
            staged.append(k)
            nbytes += dst.numel() * 4
        # one async H2D copy per run of adjacent segments (a full training batch = one copy)
        idx = sorted(INPUT_ORDER.index(k) for k in staged)
        i = 0
        while i < len(idx):
            j = i
            while j + 1 < len(idx) and idx[j + 1] == idx[j] + 1:
                j += 1
            lo = self._seg[INPUT_ORDER[idx[i]]][0]
            o, n = self._seg[INPUT_ORDER[idx[j]]]
            self.inp_all[lo:o + n].copy_(pin[lo:o + n], non_blocking=True)
            i = j + 1
        if pin is not None:
            ev = self._pin_events[slot] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._pin_events[slot] = ev
        return nbytes

    def set_rays(self, rays) -> None:
        """Explicit-rays entry: `rays` = (B, 2, S, S, 6) [pos xyz | dir xyz] for camera 1 / camera 2, e.g. the
        output of the reference's own v3d.Camera(...).rays() (model/xunet.py:159-161,166-168); None restores the
        library's ray generation from R, t, K and cfg.ray_convention."""
        if rays is None:
            self.rays = None
            self.cbatch.rays = None
            return
        r = torch.as_tensor(np.asarray(rays) if not isinstance(rays, torch.Tensor) else rays).to(torch.float32)
        if tuple(r.shape) != (self.B, 2, self.S, self.S, 6):
            raise ValueError(f'rays has shape {tuple(r.shape)}, expected {(self.B, 2, self.S, self.S, 6)}')
        self.rays = r.to(self.device).contiguous()
        self.cbatch.rays = self.rays.data_ptr()

    def set_bucket_callback(self, fn, min_bucket_bytes: int = 0) -> int:
        """Installs fn(elem_offset, n_elems), called on the host during backward() whenever a contiguous range of the flat
        gradient buffer is final (see xunet_set_grad_bucket_callback).  fn=None removes it.  Returns the bucket count."""
        if fn is None:
            _lib.check(self.lib.xunet_set_grad_bucket_callback(self.h, _lib.BUCKET_FN(0), None, 0), 'bucket_callback')
            self._bucket_cb = None
            return 0
        cb = _lib.BUCKET_FN(lambda user, off, n: fn(int(off), int(n)))
        _lib.check(self.lib.xunet_set_grad_bucket_callback(self.h, cb, None, int(min_bucket_bytes)), 'bucket_callback')
        self._bucket_cb = cb
        return int(self.lib.xunet_grad_bucket_count(self.h))

    def forward(self, flat_params: torch.Tensor, *, train: bool, seed: Optional[int] = None) -> torch.Tensor:
        assert flat_params.dtype == torch.float32 and flat_params.is_cuda and flat_params.numel() == self.nparams
        if seed is not None:
            self.seed.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.xunet_forward(self.h, flat_params.data_ptr(), C.byref(self.cbatch), int(train),
                                          self.seed.data_ptr(), self.ws.data_ptr(), self.eps.data_ptr(), st), 'xunet_forward')
        return self.eps

    def backward(self, flat_params: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """loss = ||eps - noise||_F and its parameter gradient (train.py:62-71); follows forward()."""
        if not self.training:
            raise RuntimeError('engine was built with training=False')
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.xunet_backward(self.h, flat_params.data_ptr(), C.byref(self.cbatch), self.inp['noise'].data_ptr(),
                                           self.seed.data_ptr(), self.ws.data_ptr(), self.grads.data_ptr(),
                                           self.loss.data_ptr(), st), 'xunet_backward')
        return self.loss, self.grads

    def count_kernels(self, flat_params: torch.Tensor) -> Tuple[int, int]:
        """(forward, backward) kernel launches of this plan, counted from a throw-away graph capture."""
        nf, nb = C.c_int(0), C.c_int(0)
        g = self.grads.data_ptr() if self.training else None
        _lib.check(self.lib.xunet_count_kernels(self.h, flat_params.data_ptr(), C.byref(self.cbatch), self.inp['noise'].data_ptr(),
                                                self.seed.data_ptr(), self.ws.data_ptr(), g, self.loss.data_ptr(),
                                                C.byref(nf), C.byref(nb)), 'xunet_count_kernels')
        return nf.value, nb.value

    # ---- introspection for parity tests ---------------------------------------------------------------
    def taps(self) -> Dict[str, Tuple[Tuple[int, ...], int, bool, int]]:
        if self._taps is None:
            out = OrderedDict()
            name, dims, off, isf, goff = C.c_char_p(), (C.c_int * 4)(), C.c_longlong(), C.c_int(), C.c_longlong()
            for i in range(self.lib.xunet_tap_count(self.h)):
                _lib.check(self.lib.xunet_tap(self.h, i, C.byref(name), C.byref(dims), C.byref(off), C.byref(isf), C.byref(goff)), 'tap')
                out[name.value.decode()] = (tuple(dims), int(off.value), bool(isf.value), int(goff.value))
            self._taps = out
        return self._taps

    def read_tap(self, name: str, grad: bool = False) -> torch.Tensor:
        dims, off, isf, goff = self.taps()[name]
        if grad:
            if goff < 0:
                raise KeyError(f'{name} has no gradient')
            off = goff
        dt = torch.float32 if isf else self.act_dtype
        n = int(np.prod(dims))
        raw = self.ws[off: off + n * (4 if dt == torch.float32 else 2)]
        return raw.view(dt).reshape(dims).to(torch.float32).clone()


class XUNet:
    """Drop-in for the reference's `XUNet` Flax module as used by train.py:39-43,63-66 and sampling.py:64,100-102,131-132."""

    def __init__(self, ch=32, ch_mult=(1, 2), emb_ch=32, num_res_blocks=2, attn_resolutions=(8, 16, 32), attn_heads=4,
                 dropout=0.1, use_pos_emb=False, use_ref_pose_emb=False, *, dtype='bf16', ray_convention='v3d130_ij',
                 device=None):
        self.config = XUNetConfig(ch=ch, ch_mult=tuple(ch_mult), emb_ch=emb_ch, num_res_blocks=num_res_blocks,
                                  attn_resolutions=tuple(attn_resolutions), attn_heads=attn_heads, dropout=dropout,
                                  use_pos_emb=use_pos_emb, use_ref_pose_emb=use_ref_pose_emb, dtype=dtype,
                                  ray_convention=ray_convention)
        self.device = device
        self._engines: Dict[Tuple[int, int, bool], Engine] = {}

    @classmethod
    def from_config(cls, cfg: XUNetConfig, device=None) -> "XUNet":
        m = cls.__new__(cls)
        m.config, m.device, m._engines = cfg, device, {}
        return m

    def engine(self, B: int, S: int, training: bool = False) -> Engine:
        key = (B, S, bool(training))
        if key not in self._engines:
            self._engines[key] = Engine(self.config, B, S, training, self.device)
        return self._engines[key]

    # ---- parameters -------------------------------------------------------------------------------------
    def param_spec(self, S: int, B: int = 1) -> "OrderedDict[str, Tuple[Tuple[int, ...], int]]":
        return self.engine(B, S, False).spec

    def tree_from_flat(self, flat: torch.Tensor, S: int, B: int = 1) -> ParamTree:
        spec = self.param_spec(S, B)
        leaves = {name: flat[off: off + int(np.prod(shape))].view(shape) for name, (shape, off) in spec.items()}
        tree = ParamTree(_nest(leaves))
        tree.flat, tree.spec = flat, spec
        return tree

    def flat_from_tree(self, params, S: int, B: int = 1) -> torch.Tensor:
        """Packs any nested dict of arrays in the Flax tree shape (e.g. an imported checkpoint) into the flat buffer."""
        if isinstance(params, ParamTree) and params.flat is not None:
            return params.flat
        eng = self.engine(B, S, False)
        leaves = _flatten(params)
        missing = set(eng.spec) - set(leaves)
        extra = set(leaves) - set(eng.spec)
        if missing or extra:
            raise KeyError(f'parameter tree mismatch: missing {sorted(missing)[:3]}..., unexpected {sorted(extra)[:3]}...')
        host = torch.empty(eng.nparams, dtype=torch.float32)
        for name, (shape, off) in eng.spec.items():
            v = torch.as_tensor(np.asarray(leaves[name]) if not isinstance(leaves[name], torch.Tensor) else leaves[name])
            if tuple(v.shape) != tuple(shape):
                raise ValueError(f'{name}: shape {tuple(v.shape)} != {tuple(shape)}')
            host[off: off + v.numel()] = v.reshape(-1).to(torch.float32).cpu()
        return host.to(eng.device)

    def init(self, rngs, batch, *, cond_mask=None, train=True, zero_init=True, on_device=False) -> Dict[str, ParamTree]:
        """Flax-style initialisation (train.py:41-43): lecun_normal kernels, zero biases, GroupNorm scale 1,
        zero-initialised Conv_1 kernels (out_init_scale, model/xunet.py:11-12).  The parameter set depends on the
        image side of `batch['x']` (which levels get attention).  on_device=True draws the same distributions with the
        CUDA generator straight into the flat device buffer (seconds instead of tens of seconds for the 439 M-parameter
        model; a different random stream than the host path)."""
        x = batch['x']
        B, S = int(x.shape[0]), int(x.shape[1])
        eng = self.engine(B, S, False)
        seed = _seed_of(rngs.get('params') if isinstance(rngs, dict) else rngs)
        gdev = eng.device if on_device else 'cpu'
        g = torch.Generator(device=gdev).manual_seed(seed)
        host = torch.zeros(eng.nparams, dtype=torch.float32, device=gdev)
        for name, (shape, off) in eng.spec.items():
            leaf = name.rsplit('/', 1)[-1]
            n = int(np.prod(shape))
            if leaf == 'kernel':
                if zero_init and is_zero_init_kernel(name):
                    continue
                fan_in = shape[1] * shape[2] * shape[3] if len(shape) == 5 else shape[0]
                v = host[off: off + n]
                torch.nn.init.trunc_normal_(v, mean=0., std=1., a=-2., b=2., generator=g)
                v.mul_(math.sqrt(1.0 / fan_in) / 0.87962566103423978)
            elif leaf == 'scale':
                host[off: off + n] = 1.0
            elif leaf == 'bias':
                pass
            else:  # pos_emb, ref_pose_emb_*: normal(stddev=1/sqrt(D))  model/xunet.py:182-191
                host[off: off + n] = torch.randn(n, generator=g, device=gdev) / math.sqrt(POSE_EMB_DIM)
        return {'params': self.tree_from_flat(host.to(eng.device), S, B)}

    # ---- forward ------------------------------------------------------------------------------------------
    def apply(self, variables, batch, *, cond_mask, train: bool, rngs=None, rays=None) -> torch.Tensor:
        """XUNet.apply({'params': p}, batch, cond_mask=, train=, rngs={'dropout': key}) -> eps_hat (B,S,S,3), fp32, on device.
        rays (optional, also accepted as batch['rays']): precomputed (B,2,S,S,6) camera rays, see Engine.set_rays."""
        x = batch['x']
        B, S = int(x.shape[0]), int(x.shape[1])
        if tuple(np.shape(cond_mask)) != (B,):
            raise AssertionError(f'cond_mask.shape == (B,) violated: {np.shape(cond_mask)}')   # model/xunet.py:176
        eng = self.engine(B, S, False)
        flat = self.flat_from_tree(variables['params'], S, B)
        eng.load_inputs(batch, cond_mask=cond_mask)
        eng.set_rays(rays if rays is not None else batch.get('rays'))
        seed = _seed_of(rngs.get('dropout') if isinstance(rngs, dict) else rngs)
        try:
            return eng.forward(flat, train=train, seed=seed).clone()
        finally:
            eng.set_rays(None)
