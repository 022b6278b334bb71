"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/xunet_b200.h declares,
and its static plan (parameter tree, workspace) agrees with the oracle's walk of model/xunet.py.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import xunet_ref as R
from novel_view_synthesis_3d_b200 import _lib, XUNetConfig, SMALL, FULL_3DIM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'xunet_b200.h')).read()
    declared = set(re.findall(r'\b(xunet_[a-z_0-9]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name)
    assert lib.xunet_version() >= 100


def _spec(lib, cfg: XUNetConfig, B, S, training=0, dtype=1):
    h = C.c_void_p()
    cs = cfg.c_struct()
    assert lib.xunet_create(C.byref(cs), B, S, dtype, training, C.byref(h)) == 0, lib.xunet_last_error()
    spec = {}
    name, ndim, shape, off = C.c_char_p(), C.c_int(), (C.c_longlong * 5)(), C.c_longlong()
    for i in range(lib.xunet_param_leaves(h)):
        assert lib.xunet_param_leaf(h, i, C.byref(name), C.byref(ndim), C.byref(shape), C.byref(off)) == 0
        spec[name.value.decode()] = (tuple(shape[k] for k in range(ndim.value)), off.value)
    n, ws = lib.xunet_param_count(h), lib.xunet_workspace_bytes(h)
    lib.xunet_destroy(h)
    return spec, n, ws


@pytest.mark.parametrize('cfg,S', [
    (SMALL, 64), (SMALL, 128), (FULL_3DIM, 128), (FULL_3DIM, 64),
    (XUNetConfig(ch=32, ch_mult=(1, 2), emb_ch=32, num_res_blocks=1, attn_resolutions=(8, 16), attn_heads=2,
                 use_pos_emb=True, use_ref_pose_emb=True), 16),
    (XUNetConfig(ch=64, ch_mult=(1, 2, 4), emb_ch=64, num_res_blocks=1, attn_resolutions=(8,), attn_heads=4), 32),
])
def test_param_tree_matches_flax_walk(lib, cfg, S):
    from tests.util import to_ref_cfg
    spec, n, ws = _spec(lib, cfg, 2, S)
    ref = R.param_shapes(to_ref_cfg(cfg), S)
    assert set(spec) == set(ref)
    for k, shp in ref.items():
        assert spec[k][0] == tuple(shp), k
    assert n == sum(int(np.prod(s)) for s in ref.values())
    # leaves tile the flat buffer exactly
    spans = sorted((off, off + int(np.prod(shp))) for shp, off in spec.values())
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert ws > 0


def test_known_param_counts(lib):
    assert _spec(lib, SMALL, 1, 64)[1] == 1054211
    assert _spec(lib, SMALL, 1, 128)[1] == 902915
    assert _spec(lib, FULL_3DIM, 1, 128)[1] == 438831363
    assert _spec(lib, FULL_3DIM, 1, 64)[1] == 449877251


def test_create_rejects_bad_configs(lib):
    h = C.c_void_p()
    bad = [XUNetConfig(ch=48), XUNetConfig(emb_ch=30), XUNetConfig(dropout=1.0),
           XUNetConfig(ch=32, ch_mult=(1, 2), attn_heads=3)]
    for cfg in bad:
        cs = cfg.c_struct()
        assert lib.xunet_create(C.byref(cs), 2, 64, 1, 0, C.byref(h)) != 0
        assert lib.xunet_last_error()
    cs = SMALL.c_struct()
    assert lib.xunet_create(C.byref(cs), 2, 63, 1, 0, C.byref(h)) != 0      # side not divisible by 2^(L-1)
    assert lib.xunet_create(C.byref(cs), 2, 64, 7, 0, C.byref(h)) != 0      # bad dtype


def test_training_workspace_is_larger(lib):
    assert _spec(lib, SMALL, 2, 64, training=1)[2] > _spec(lib, SMALL, 2, 64, training=0)[2]
    assert _spec(lib, SMALL, 2, 64, dtype=0)[2] > _spec(lib, SMALL, 2, 64, dtype=1)[2]


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU (no eager / CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from novel_view_synthesis_3d_b200 import XUNet
    with pytest.raises(RuntimeError, match='CUDA device'):
        XUNet().engine(1, 64)


def test_library_is_built_from_the_sources_in_the_tree():
    """A stale or failed build must not go unnoticed (the .so is git-ignored and travels with the tree): build() is a no-op when
    the stamp matches the sources, recompiles otherwise, and raises on any nvcc error."""
    from novel_view_synthesis_3d_b200 import build as B
    lib = B.build()
    assert lib.exists()
    deps = list(B.CSRC.glob('*.cu')) + list(B.CSRC.glob('*.cuh')) + list(B.CSRC.glob('*.h')) + list(B.INCLUDE.glob('*.h'))
    assert (B.OBJ_DIR / 'stamp').read_text() == B._digest(deps)
