"""SRN reader (dataset/data_loader.py, data_util.py, util.py of the reference) on a synthetic scene tree. CPU only."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')
from novel_view_synthesis_3d_b200 import srn_data as D


def _make_tree(root, n_inst=2, n_views=3, H=40, W=48):
    rng = np.random.RandomState(0)
    for i in range(n_inst):
        d = os.path.join(root, f'obj{i}')
        os.makedirs(os.path.join(d, 'rgb')); os.makedirs(os.path.join(d, 'pose'))
        with open(os.path.join(d, 'intrinsics.txt'), 'w') as fh:
            fh.write('131.25 64.0 64.0 0.\n0. 0. 0.\n1.\n128 128\n')
        for v in range(n_views):
            img = rng.randint(0, 256, (H, W, 3), dtype=np.uint8)
            img[..., 0] = 10 * (v + 1)                       # B channel constant -> checks BGR->RGB
            cv2.imwrite(os.path.join(d, 'rgb', f'{v:06d}.png'), img)
            pose = np.eye(4); pose[:3, 3] = [i, v, 1.3]
            fmt = ' '.join(f'{x:.6f}' for x in pose.reshape(-1)) if v % 2 == 0 else '\n'.join(' '.join(f'{x:.6f}' for x in r) for r in pose)
            open(os.path.join(d, 'pose', f'{v:06d}.txt'), 'w').write(fmt + '\n')


def test_intrinsics_rescale_and_pose_formats(tmp_path):
    _make_tree(str(tmp_path))
    K = D.parse_intrinsics(str(tmp_path / 'obj0' / 'intrinsics.txt'), 64)
    assert np.allclose(K, [[65.625, 0, 32], [0, 65.625, 32], [0, 0, 1]])       # f*S/height, cx/width*S (util.py:64-67)
    p0 = D.load_pose(str(tmp_path / 'obj1' / 'pose' / '000000.txt'))           # single-line format
    p1 = D.load_pose(str(tmp_path / 'obj1' / 'pose' / '000001.txt'))           # four-line format
    assert p0.shape == (4, 4) and np.allclose(p0[:3, 3], [1, 0, 1.3]) and np.allclose(p1[:3, 3], [1, 1, 1.3])


def test_load_rgb_crop_resize_range(tmp_path):
    _make_tree(str(tmp_path))
    img = D.load_rgb(str(tmp_path / 'obj0' / 'rgb' / '000001.png'), 16)
    assert img.shape == (16, 16, 3) and img.dtype == np.float32 and -1.0 <= img.min() and img.max() <= 1.0
    assert np.allclose(img[..., 2], 20 / 255 * 2 - 1, atol=1e-6)               # constant channel written as B comes back as img[...,2]
    full = D.load_rgb(str(tmp_path / 'obj0' / 'rgb' / '000001.png'))
    assert full.shape == (40, 40, 3)                                           # centre square crop of 40x48


def test_batches_follow_the_reference_contract(tmp_path):
    _make_tree(str(tmp_path))
    ds = D.SRNScenes(str(tmp_path), img_sidelength=16, host_diffusion=True, seed=1)
    assert len(ds) == 6
    b = next(ds.batches(4, shuffle=True))
    assert b['x'].shape == (4, 16, 16, 3) and b['target'].shape == (4, 16, 16, 3) and b['z'].shape == (4, 16, 16, 3)
    assert b['R1'].shape == (4, 3, 3) and b['t1'].shape == (4, 3) and b['K'].shape == (4, 3, 3) and b['logsnr'].shape == (4,)
    assert b['z'].dtype == np.float64 and b['noise'].dtype == np.float64 and b['x'].dtype == np.float32   # data_loader.py:102-112
    assert np.all(b['t1'][:, 0] == b['t2'][:, 0])                             # source and target come from the same instance
    with pytest.raises(AssertionError):
        D.SRNScenes(str(tmp_path / 'empty'))
