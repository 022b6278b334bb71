"""Flax msgpack checkpoint format (train.py:159-167 / sampling.py:106-114): byte-level known answer + round trips. CPU only."""
import os

import msgpack
import numpy as np
import torch

from oracle import xunet_ref as R
from novel_view_synthesis_3d_b200 import checkpoint as ck


def test_ndarray_ext_encoding_known_answer():
    # flax.serialization: ExtType(1, packb((shape, dtype.name, bytes)))
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    blob = ck.msgpack_serialize({'w': a})
    raw = msgpack.unpackb(blob, raw=False)
    assert isinstance(raw['w'], msgpack.ExtType) and raw['w'].code == 1
    shape, dtype, buf = msgpack.unpackb(raw['w'].data, raw=False)
    assert tuple(shape) == (2, 3) and dtype == 'float32' and buf == a.tobytes()
    back = ck.msgpack_restore(blob)
    assert back['w'].dtype == np.float32 and np.array_equal(back['w'], a)


def test_roundtrip_param_tree_with_device_axis(tmp_path):
    params = R.init_params(R.SMALL, 64, seed=1, zero_init=False, dtype=torch.float32)
    path = ck.save_checkpoint(str(tmp_path), params, step=0, prefix='model', add_device_axis=True)
    assert os.path.basename(path) == 'model0'                    # the name sampling.py:109 looks for
    raw = ck.msgpack_restore(open(path, 'rb').read())
    assert raw['Conv_0']['kernel'].shape == (1, 1, 3, 3, 3, 32)   # leading pmap axis as in the reference's files
    tree = ck.restore_checkpoint(str(tmp_path), prefix='model')
    flat_a, flat_b = R.flatten(params), R.flatten(tree)
    assert set(flat_a) == set(flat_b)
    for k in flat_a:
        assert np.array_equal(flat_a[k].numpy(), flat_b[k]), k
    # overwrite=True keeps a single file per prefix, latest wins
    ck.save_checkpoint(str(tmp_path), params, step=1000, prefix='model', add_device_axis=True)
    assert sorted(os.listdir(tmp_path)) == ['model1000']
    assert ck.restore_checkpoint(str(tmp_path / 'nope')) is None


def test_chunked_large_array_roundtrip(monkeypatch):
    monkeypatch.setattr(ck, '_MAX_CHUNK', 1024)
    a = np.random.RandomState(0).randn(40, 33).astype(np.float32)
    back = ck.msgpack_restore(ck.msgpack_serialize({'big': a, 's': np.float32(2.5)}))
    assert np.array_equal(back['big'], a) and back['s'] == np.float32(2.5)
