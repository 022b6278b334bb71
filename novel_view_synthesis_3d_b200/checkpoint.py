"""Flax checkpoint import / export for the X-UNet parameter tree  (SURVEY 8(f) row 1).

The reference saves `train_state.params` with `flax.training.checkpoints.save_checkpoint(ckpt_dir, target, step,
prefix='model', overwrite=True)` (train.py:159-167) -- every leaf carries a leading per-device axis because the state is
pmapped -- and restores it in sampling.py:106-114.  Flax (0.6.4) is not installable here, so this module restates its
published on-disk format (flax/serialization.py): a msgpack map of the nested state dict whose ndarray leaves are msgpack
ExtType(code=1) payloads = msgpack-packed `(shape, dtype_name, raw C-order bytes)`; numpy scalars are ExtType(3);
arrays larger than 2**30 bytes are stored as a dict of chunks {'__msgpack_chunked_array__': True, 'shape': {'0':..}, 'chunks': {'0':..}}.
File name: f'{prefix}{step}' inside ckpt_dir.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional

import msgpack
import numpy as np

_EXT_NDARRAY, _EXT_NATIVE_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_MAX_CHUNK = 2 ** 30


def _ndarray_to_bytes(arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    return msgpack.packb((arr.shape, arr.dtype.name, arr.tobytes('C')), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
    return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(_EXT_NATIVE_COMPLEX, msgpack.packb((x.real, x.imag)))
    return x


def _ext_unpack(code, data):
    if code == _EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == _EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    if code == _EXT_NATIVE_COMPLEX:
        re_, im_ = msgpack.unpackb(data)
        return complex(re_, im_)
    return msgpack.ExtType(code, data)


def _chunk(tree):
    if isinstance(tree, dict):
        return {k: _chunk(v) for k, v in tree.items()}
    if isinstance(tree, np.ndarray) and tree.size * tree.dtype.itemsize > _MAX_CHUNK:
        flat = tree.reshape(-1)
        per = max(1, _MAX_CHUNK // tree.dtype.itemsize)
        chunks = [flat[i:i + per] for i in range(0, flat.size, per)]
        # flax stores tuples as {'0': .., '1': ..} dicts (msgpack strict_types)
        return {'__msgpack_chunked_array__': True, 'shape': {str(i): int(v) for i, v in enumerate(tree.shape)},
                'chunks': {str(i): c for i, c in enumerate(chunks)}}
    return tree


def _unchunk(tree):
    if isinstance(tree, dict):
        if tree.get('__msgpack_chunked_array__'):
            chunks = tree['chunks']
            parts = [chunks[str(i)] for i in range(len(chunks))]
            shp = tree['shape']
            shape = tuple(shp[str(i)] for i in range(len(shp))) if isinstance(shp, dict) else tuple(shp)
            return np.concatenate([np.asarray(p).reshape(-1) for p in parts]).reshape(shape)
        return {k: _unchunk(v) for k, v in tree.items()}
    return tree


def msgpack_serialize(tree: dict) -> bytes:
    """flax.serialization.msgpack_serialize for nested dicts of numpy arrays."""
    def to_np(t):
        if isinstance(t, dict):
            return {str(k): to_np(v) for k, v in t.items()}
        if hasattr(t, 'detach'):          # torch tensor
            return t.detach().cpu().numpy()
        return np.asarray(t) if not isinstance(t, (int, float, str, bool)) else t
    return msgpack.packb(_chunk(to_np(tree)), default=_ext_pack, strict_types=True, use_bin_type=True)


def msgpack_restore(data: bytes) -> dict:
    """flax.serialization.msgpack_restore."""
    return _unchunk(msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


def strip_device_axis(tree: dict, device_index: int = 0) -> dict:
    """The reference checkpoints pmapped params: every leaf is (n_devices, ...) (train.py:161-167). Take one replica."""
    if isinstance(tree, dict):
        return {k: strip_device_axis(v, device_index) for k, v in tree.items()}
    return np.asarray(tree)[device_index]


def save_checkpoint(ckpt_dir: str, target, step: int, prefix: str = 'model', overwrite: bool = True,
                    add_device_axis: bool = False) -> str:
    """checkpoints.save_checkpoint look-alike (train.py:161-167).  `target` = nested param dict (ParamTree / numpy / torch)."""
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f'{prefix}{step}')
    if os.path.exists(path) and not overwrite:
        raise FileExistsError(path)
    tree = target
    if add_device_axis:
        def add(t):
            if isinstance(t, dict):
                return {k: add(v) for k, v in t.items()}
            a = t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)
            return a[None]
        tree = add(target)
    if overwrite:   # flax removes older checkpoints with the same prefix when overwrite=True
        for f in os.listdir(ckpt_dir):
            if re.fullmatch(re.escape(prefix) + r'\d+', f) and f != os.path.basename(path):
                os.remove(os.path.join(ckpt_dir, f))
    tmp = path + '.tmp'
    with open(tmp, 'wb') as fh:
        fh.write(msgpack_serialize(tree))
    os.replace(tmp, path)
    return path


def latest_checkpoint(ckpt_dir: str, prefix: str = 'model') -> Optional[str]:
    best, best_step = None, -1
    if not os.path.isdir(ckpt_dir):
        return None
    for f in os.listdir(ckpt_dir):
        # flax globs f'{prefix}*' and sorts naturally: 'model0' is found both by prefix 'model' (step 0) and by prefix
        # 'model0' (sampling.py:109), where nothing follows the prefix
        m = re.fullmatch(re.escape(prefix) + r'(\d*)', f)
        if m:
            step = int(m.group(1)) if m.group(1) else 0
            if step > best_step:
                best, best_step = os.path.join(ckpt_dir, f), step
    return best


def restore_checkpoint(ckpt_dir: str, prefix: str = 'model', device_index: Optional[int] = None) -> Optional[dict]:
    """checkpoints.restore_checkpoint look-alike (sampling.py:106-110): returns the nested param dict of the latest
    f'{prefix}<step>' file, or None if there is none (the reference raises FileNotFoundError in that case, :111-112).
    If the leaves carry the pmap device axis (auto-detected on GroupNorm scales: 2-D instead of 1-D) one replica is taken."""
    path = ckpt_dir if os.path.isfile(ckpt_dir) else latest_checkpoint(ckpt_dir, prefix)
    if path is None:
        return None
    with open(path, 'rb') as fh:
        tree = msgpack_restore(fh.read())
    if 'params' in tree and isinstance(tree['params'], dict) and 'Conv_0' not in tree:
        tree = tree['params']      # a whole TrainState was saved
    probe = tree.get('GroupNorm_0', {}).get('GroupNorm_0', {}).get('scale')
    has_axis = probe is not None and np.asarray(probe).ndim == 2
    if device_index is not None or has_axis:
        tree = strip_device_axis(tree, device_index or 0)
    return tree
