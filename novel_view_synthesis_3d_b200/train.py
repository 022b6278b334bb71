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

import os

import ctypes as C
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
import torch

from . import _lib
from . import dist as xdist
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

    def apply_gradients(self, *, grads, copy: bool = False) -> "TrainState":
        """TrainState.apply_gradients (train.py:76): one fused Adam pass over the flat buffers.

        Ownership: by default the flat parameter / moment buffers are updated IN PLACE and the returned state shares
        them, i.e. the OLD state's `params` alias the new values (the reference rebinds `state = update_model(state, ..)`
        and never looks at the old one, train.py:153).  copy=True gives the Flax contract literally: the old state keeps
        its values and the new state owns fresh buffers (3 x 4 bytes per parameter of extra traffic)."""
        lib = _lib.load()
        flat_g = grads.flat if isinstance(grads, ParamTree) else grads
        if copy:
            newp = self.model.tree_from_flat(self.params.flat.clone(), self.img_sidelength, self.batch_size)
            self = replace(self, params=newp, opt_state=replace(self.opt_state, mu=self.opt_state.mu.clone(),
                                                                nu=self.opt_state.nu.clone()))
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
                       zero_init: bool = True, reference_quirks: bool = False, init_on_device: bool = False) -> TrainState:
    """train.py:36-47.  Under torch.distributed the parameters of rank 0 are broadcast (the reference gives every
    device a different init, train.py:122-123 -- an ensemble, not data parallelism)."""
    model = model or XUNet()
    sample = create_sample_data(train_batch_size, img_sidelength)
    params = model.init({'params': rng, 'dropout': rng_dropout}, sample, cond_mask=np.zeros(train_batch_size), train=True,
                        zero_init=zero_init, on_device=init_on_device)['params']
    world = xdist.world_size()
    xdist.broadcast_params(params.flat, src=0)
    opt = Adam(learning_rate)
    st = AdamState(mu=torch.zeros_like(params.flat), nu=torch.zeros_like(params.flat), count=0)
    mask_seed = int(np.asarray(rng_dropout).reshape(-1)[-1]) if rng_dropout is not None else 0
    # data parallel: every rank draws its OWN cond_mask stream (and dropout seeds, _dropout_seed) for its shard --
    # identical streams would drop sample b of every shard together, which is not a global batch of world*B
    mask_seed ^= (xdist.rank() * 0x9E3779B9) & 0x7FFFFFFF
    return TrainState(step=0, apply_fn=model.apply, params=params, tx=opt, opt_state=st, model=model,
                      batch_size=train_batch_size, img_sidelength=img_sidelength, grad_scale=1.0 / world,
                      reference_quirks=reference_quirks, _mask_rng=np.random.RandomState(mask_seed & 0x7FFFFFFF))


def _dropout_seed(step_plus_1: int) -> int:
    """Dropout seed of optimisation step `step_plus_1` (1-based) on this rank: the rank sits in the high bits."""
    return (xdist.rank() << 40) + int(step_plus_1)


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
    seed = 0 if state.reference_quirks else _dropout_seed(state.step + 1)   # PRNGKey(0) frozen at trace time, train.py:66
    eng.forward(state.params.flat, train=True, seed=seed)
    loss, grads = eng.backward(state.params.flat)
    xdist.allreduce_sum_(grads, bucket_elems=64 << 20)
    # the loss is a copy; `grads` are views of the engine's gradient bucket (valid until the next backward of this plan)
    return loss[0].clone(), state.model.tree_from_flat(grads, S, B)


def update_model(state: TrainState, grads) -> TrainState:
    """train.py:74-76."""
    return state.apply_gradients(grads=grads)


class TrainStep:
    """The fused production step: pinned H2D staging -> forward -> backward (gradient buckets all-reduced over NCCL as they
    become final, overlapping the rest of the backward) -> Adam with 1/world folded in.  Semantically apply_model +
    update_model (train.py:142-153).

    Execution modes (chosen at the first call):
      world == 1                : CUDA graph [forward + backward + Adam]
      world > 1, NCCL           : ONE CUDA graph [forward + backward + bucketed all-reduces + Adam]; the collectives are
                                  captured on NCCL's stream as parallel branches of the graph
      world > 1, other backends : graph [forward + backward]; eager bucketed all-reduce; graph [Adam]   (gloo in tests, or
                                  XUNET_DP_EAGER_COLLECTIVES=1, or when capturing the collectives fails)
      use_graph=False           : everything eager; the bucket hook still overlaps the all-reduces with the backward

    A captured graph holds the NCCL kernels of the process group it was captured with: drop the TrainStep (or call
    `release_graphs()`) before `torch.distributed.destroy_process_group()`.
    """

    def __init__(self, state: TrainState, *, use_graph: bool = True, bucket_mb: float = 128.0, allreduce: bool = True):
        import os
        self.state = state
        self.eng: Engine = state.model.engine(state.batch_size, state.img_sidelength, True)
        self.lib = _lib.load()
        self.dev = self.eng.device
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.world = xdist.world_size()
        self.use_graph = use_graph
        self.graph_fb = None      # forward+backward (+ all-reduce + adam when everything is in one graph)
        self.graph_opt = None     # adam (two-graph mode only)
        self.stream = torch.cuda.Stream(device=self.dev) if use_graph else None
        self.launches_per_step = None
        self.allreduce = allreduce and self.world > 1       # allreduce=False: measurement knob (compute-only step at N>1)
        # XUNET_DP_BUCKET_MB overrides the bucket size (experiments: smaller buckets start the collectives earlier but add launches)
        if os.environ.get('XUNET_DP_BUCKET_MB'):
            bucket_mb = float(os.environ['XUNET_DP_BUCKET_MB'])
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.reducer = xdist.GradReducer(self.eng.grads) if self.world > 1 else None
        if self.reducer is not None:
            self.reducer.enabled = self.allreduce
        backend = torch.distributed.get_backend() if self.world > 1 else None
        self.capture_collectives = (self.world > 1 and backend == 'nccl' and use_graph and
                                    os.environ.get('XUNET_DP_EAGER_COLLECTIVES') != '1')
        self.mode = None
        self._synced_count = None
        self._sync_counters()

    def release_graphs(self):
        self.graph_fb = self.graph_opt = None

    def _sync_counters(self):
        """Adam step counter and dropout seed live on the device (a replayed graph sees them advance); re-derive them from
        the host-side state whenever something else (apply_model / update_model, a restored checkpoint) moved it."""
        s = self.state
        if self._synced_count != s.opt_state.count:
            self.step_dev.fill_(s.opt_state.count)
            self.eng.seed.fill_(0 if s.reference_quirks else _dropout_seed(s.opt_state.count))
            self._synced_count = s.opt_state.count

    def _fwd_bwd(self):
        e, s = self.eng, self.state
        if not s.reference_quirks:
            e.seed.add_(1)
        self.step_dev.add_(1)
        e.forward(s.params.flat, train=True)
        if self.reducer is not None:
            self.reducer.begin()
        e.backward(s.params.flat)

    def _adam(self):
        e, s = self.eng, self.state
        st = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.lib.xunet_adam_step(s.params.flat.data_ptr(), e.grads.data_ptr(), s.opt_state.mu.data_ptr(),
                                            s.opt_state.nu.data_ptr(), s.params.flat.numel(), 0, self.step_dev.data_ptr(),
                                            s.tx.learning_rate, s.tx.b1, s.tx.b2, s.tx.eps, 1.0 / self.world, st), 'adam')

    def _full_step(self):
        self._fwd_bwd()
        if self.reducer is not None:
            self.reducer.wait()
        self._adam()

    def _set_hook(self, on: bool):
        if self.reducer is None:
            return
        self.eng.set_bucket_callback(self.reducer.on_bucket if on else None, self.bucket_bytes)

    def _capture(self):
        torch.cuda.synchronize(self.dev)
        s = self.state
        seed0, step0 = self.eng.seed.clone(), self.step_dev.clone()
        backup = None
        one_graph = self.world == 1 or self.capture_collectives
        if one_graph:
            # warm-up and capture run real Adam updates: snapshot params/moments and restore them afterwards
            backup = (s.params.flat.clone(), s.opt_state.mu.clone(), s.opt_state.nu.clone())
        try:
            with torch.cuda.stream(self.stream):
                self._set_hook(one_graph)
                if one_graph:
                    self._full_step()                    # warm-up outside capture (also initialises the NCCL communicator)
                else:
                    self._fwd_bwd()
                self.stream.synchronize()
                self.graph_fb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_fb, stream=self.stream):
                    if one_graph:
                        self._full_step()
                    else:
                        self._fwd_bwd()
                if not one_graph:
                    self.graph_opt = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_opt, stream=self.stream):
                        self._adam()
            self.mode = 'one_graph' if one_graph else 'two_graphs_eager_collectives'
        except Exception:
            if not (one_graph and self.world > 1):
                raise
            # capturing the collectives failed: fall back to eager all-reduces between two graphs
            torch.cuda.synchronize(self.dev)
            self.capture_collectives = False
            self.graph_fb = self.graph_opt = None
            if backup is not None:
                s.params.flat.copy_(backup[0]); s.opt_state.mu.copy_(backup[1]); s.opt_state.nu.copy_(backup[2])
            self.eng.seed.copy_(seed0)
            self.step_dev.copy_(step0)
            return self._capture()
        finally:
            self._set_hook(False)
        torch.cuda.synchronize(self.dev)
        if backup is not None:
            s.params.flat.copy_(backup[0]); s.opt_state.mu.copy_(backup[1]); s.opt_state.nu.copy_(backup[2])
        self.eng.seed.copy_(seed0)
        self.step_dev.copy_(step0)

    def __call__(self, batch: dict, noise, cond_mask=None) -> torch.Tensor:
        """One optimisation step on host (or device) inputs; returns the loss as a 0-d device tensor (a view of the engine's
        loss slot: read or copy it before the next step)."""
        s = self.state
        self._sync_counters()
        mask = _cond_mask(s, self.eng.B) if cond_mask is None else cond_mask
        self.h2d_bytes = self.eng.load_inputs(batch, cond_mask=mask, noise=noise)
        if self.use_graph:
            if self.graph_fb is None:
                self._capture()
            self.graph_fb.replay()
            if self.graph_opt is not None:
                if self.allreduce:
                    xdist.allreduce_sum_(self.eng.grads, bucket_elems=self.bucket_bytes // 4)
                self.graph_opt.replay()
        else:
            self.mode = 'eager'
            self._set_hook(True)
            try:
                self._full_step()
            finally:
                self._set_hook(False)
        s.step += 1
        s.opt_state.count += 1
        self._synced_count = s.opt_state.count
        return self.eng.loss[0]
