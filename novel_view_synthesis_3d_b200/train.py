"""Training step of the reference (train.py:36-76) on the B200 path.

Reference call shapes are kept:
    state = create_train_state(rng, rng_dropout, learning_rate, train_batch_size, img_sidelength)
    loss, grads = apply_model(state, x, z, logsnr, R1, t1, R2, t2, K, noise)
    state = update_model(state, grads)
Differences by design: the batch is SHARDED across ranks and gradients are averaged with one
NCCL all-reduce of the flat gradient bucket (the reference's pmap never calls pmean, train.py:49-76); cond_mask and the
dropout seed are fresh per step unless reference_quirks=True (the reference freezes both at trace time, train.py:64,66).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
import torch

from . import _lib
from .xunet import XUNet, XUNetConfig, ParamTree, Engine


@dataclass
class AdamState:
    """optax.adam(learning_rate) state (train.py:45): b1=.9, b2=.999, eps=1e-8."""
    mu: torch.Tensor
    nu: torch.Tensor
    count: int = 0


@dataclass
class Adam:
    learning_rate: float = 1e-4
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8


@dataclass
class TrainState:
    """Look-alike of flax.training.train_state.TrainState (fields used by train.py / sampling.py)."""
    step: int
    apply_fn: object
    params: ParamTree
    tx: Adam
    opt_state: AdamState
    model: XUNet = None
    batch_size: int = 0
    img_sidelength: int = 0
    grad_scale: float = 1.0
    reference_quirks: bool = False
    _mask_rng: np.random.RandomState = None
    _frozen_mask: Optional[np.ndarray] = None

    def apply_gradients(self, *, grads) -> "TrainState":
        """TrainState.apply_gradients (train.py:76): one fused Adam pass over the flat buffers.  Functional in
        interface (returns the new state); the flat device buffers are updated in place."""
        lib = _lib.load()
        flat_g = grads.flat if isinstance(grads, ParamTree) else grads
        p = self.params.flat
        count = self.opt_state.count + 1
        st = torch.cuda.current_stream(p.device).cuda_stream
        _lib.check(lib.xunet_adam_step(p.data_ptr(), flat_g.data_ptr(), self.opt_state.mu.data_ptr(),
                                       self.opt_state.nu.data_ptr(), p.numel(), count, None, self.tx.learning_rate,
                                       self.tx.b1, self.tx.b2, self.tx.eps, self.grad_scale, st), 'xunet_adam_step')
        return replace(self, step=self.step + 1, opt_state=replace(self.opt_state, count=count))


def create_sample_data(batch_size, img_sidelength):
    """train.py:23-34."""
    r = np.random.random
    return dict(x=r((batch_size, img_sidelength, img_sidelength, 3)), z=r((batch_size, img_sidelength, img_sidelength, 3)),
                logsnr=r((batch_size,)), R1=r((batch_size, 3, 3)), t1=r((batch_size, 3)), R2=r((batch_size, 3, 3)),
                t2=r((batch_size, 3)), K=r((batch_size, 3, 3)), noise=r((batch_size, img_sidelength, img_sidelength, 3)))


def create_train_state(rng, rng_dropout, learning_rate, train_batch_size, img_sidelength, *, model: XUNet = None,
                       zero_init: bool = True, reference_quirks: bool = False) -> TrainState:
    """train.py:36-47.  Under torch.distributed the parameters of rank 0 are broadcast (the reference gives every
    device a different init, train.py:122-123 -- an ensemble, not data parallelism)."""
    model = model or XUNet()
    sample = create_sample_data(train_batch_size, img_sidelength)
    params = model.init({'params': rng, 'dropout': rng_dropout}, sample, cond_mask=np.zeros(train_batch_size), train=True,
                        zero_init=zero_init)['params']
    world = 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        world = torch.distributed.get_world_size()
        if world > 1:
            torch.distributed.broadcast(params.flat, src=0)
    opt = Adam(learning_rate)
    st = AdamState(mu=torch.zeros_like(params.flat), nu=torch.zeros_like(params.flat), count=0)
    mask_seed = int(np.asarray(rng_dropout).reshape(-1)[-1]) if rng_dropout is not None else 0
    return TrainState(step=0, apply_fn=model.apply, params=params, tx=opt, opt_state=st, model=model,
                      batch_size=train_batch_size, img_sidelength=img_sidelength, grad_scale=1.0 / world,
                      reference_quirks=reference_quirks, _mask_rng=np.random.RandomState(mask_seed & 0x7FFFFFFF))


def _cond_mask(state: TrainState, B: int) -> np.ndarray:
    # train.py:64  np.where(np.random.random(B) > 0.1, 1, 0)
    if state.reference_quirks:
        if state._frozen_mask is None:
            state._frozen_mask = np.where(state._mask_rng.random_sample(B) > 0.1, 1, 0).astype(np.float32)
        return state._frozen_mask
    return np.where(state._mask_rng.random_sample(B) > 0.1, 1, 0).astype(np.float32)


def apply_model(state: TrainState, batch_x, batch_z, batch_logsnr, batch_R1, batch_t1, batch_R2, batch_t2, batch_K,
                batch_noise, *, cond_mask=None):
    """train.py:49-72: forward with train=True, loss ||eps_hat - noise||_F, gradients w.r.t. all parameters.
    Returns (loss: 0-d device tensor, grads: ParamTree view of the flat gradient bucket).  With
    torch.distributed initialised the gradient bucket is all-reduced (sum) here; update_model applies 1/world."""
    B, S = int(batch_x.shape[0]), int(batch_x.shape[1])
    eng = state.model.engine(B, S, True)
    batch = dict(x=batch_x, z=batch_z, logsnr=batch_logsnr, R1=batch_R1, t1=batch_t1, R2=batch_R2, t2=batch_t2, K=batch_K)
    mask = _cond_mask(state, B) if cond_mask is None else cond_mask
    eng.load_inputs(batch, cond_mask=mask, noise=batch_noise)
    seed = 0 if state.reference_quirks else state.step + 1       # PRNGKey(0) frozen at trace time, train.py:66
    eng.forward(state.params.flat, train=True, seed=seed)
    loss, grads = eng.backward(state.params.flat)
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(grads)
    return loss[0], state.model.tree_from_flat(grads, S, B)


def update_model(state: TrainState, grads) -> TrainState:
    """train.py:74-76."""
    return state.apply_gradients(grads=grads)


class TrainStep:
    """The fused production step: pinned H2D staging -> forward -> backward -> (NCCL all-reduce) -> Adam, optionally
    replayed as one CUDA graph.  Semantically apply_model + update_model."""

    def __init__(self, state: TrainState, *, use_graph: bool = True):
        self.state = state
        self.eng: Engine = state.model.engine(state.batch_size, state.img_sidelength, True)
        self.lib = _lib.load()
        self.dev = self.eng.device
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.step_dev.fill_(state.opt_state.count)
        self.world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.use_graph = use_graph
        self.graph_fb = None      # forward+backward
        self.graph_opt = None     # adam
        self.stream = torch.cuda.Stream(device=self.dev) if use_graph else None
        self.launches_per_step = None

    def _fwd_bwd(self):
        e, s = self.eng, self.state
        e.seed.add_(1)
        self.step_dev.add_(1)
        e.forward(s.params.flat, train=True)
        e.backward(s.params.flat)

    def _adam(self):
        e, s = self.eng, self.state
        st = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.lib.xunet_adam_step(s.params.flat.data_ptr(), e.grads.data_ptr(), s.opt_state.mu.data_ptr(),
                                            s.opt_state.nu.data_ptr(), s.params.flat.numel(), 0, self.step_dev.data_ptr(),
                                            s.tx.learning_rate, s.tx.b1, s.tx.b2, s.tx.eps, 1.0 / self.world, st), 'adam')

    def _capture(self):
        torch.cuda.synchronize(self.dev)
        seed0, step0 = self.eng.seed.clone(), self.step_dev.clone()
        with torch.cuda.stream(self.stream):
            self._fwd_bwd()                      # warm-up outside capture
            self.stream.synchronize()
            self.graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb, stream=self.stream):
                self._fwd_bwd()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, stream=self.stream):
                self._adam()
        torch.cuda.synchronize(self.dev)
        self.eng.seed.copy_(seed0)
        self.step_dev.copy_(step0)

    def __call__(self, batch: dict, noise, cond_mask=None) -> torch.Tensor:
        """One optimisation step on host (or device) inputs; returns the loss as a 0-d device tensor."""
        s = self.state
        mask = _cond_mask(s, self.eng.B) if cond_mask is None else cond_mask
        self.h2d_bytes = self.eng.load_inputs(batch, cond_mask=mask, noise=noise)
        if self.use_graph:
            if self.graph_fb is None:
                self._capture()
            self.graph_fb.replay()
        else:
            self._fwd_bwd()
        if self.world > 1:
            torch.distributed.all_reduce(self.eng.grads)
        if self.use_graph:
            self.graph_opt.replay()
        else:
            self._adam()
        s.step += 1
        s.opt_state.count += 1
        return self.eng.loss[0]
