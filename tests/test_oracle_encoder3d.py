"""CPU tests of the voxel-encoder oracle (oracle/encoder3d_ref.py, test infrastructure) against the definition of a
Keras Conv3D('same') and the structure of the reference's encoder3D."""
import numpy as np
import pytest
import torch

from oracle import anytime_ref as ar, encoder3d_ref as e3


def test_structure_and_mac_counts():
    st = e3.MODELNET_ENCODER                                            # test_modelnet_VAE_dr.py:172-181
    shapes = e3.weight_shapes(st)
    assert len(shapes) == 21 and shapes[0][1] == (4, 4, 4, 1, 64) and shapes[-1][1] == (4, 4, 4, 512, 128)
    assert [s for n, s in shapes if n.endswith('kernel')][1:4] == [(4, 4, 4, 64, 128), (4, 4, 4, 128, 256), (4, 4, 4, 256, 512)]
    alg, dense = e3.encoder_macs(st)
    assert dense == 4_160_749_568 and alg == 3_438_050_816               # 8.32 GFLOP / object
    assert e3.same_pads(64, 4, 2) == (1, 1) and e3.same_pads(4, 4, 1) == (1, 2) and e3.same_pads(5, 4, 2) == (1, 2)


@pytest.mark.parametrize('stride,n', [(2, 6), (1, 4), (2, 5)])
def test_conv3d_same_matches_definition(stride, n):
    rng = np.random.default_rng(stride * 10 + n)
    x = rng.standard_normal((2, n, n, n, 2))
    k = rng.standard_normal((4, 4, 4, 2, 3))
    got = e3.conv3d_same(torch.from_numpy(x).permute(0, 4, 1, 2, 3), torch.from_numpy(k), stride)
    np.testing.assert_allclose(got.permute(0, 2, 3, 4, 1).numpy(), e3.conv3d_same_definition(x, k, stride), atol=1e-12)


def test_forward_shapes_pool_and_final_activation():
    small = dict(e3.MODELNET_ENCODER, input_shape=[16, 16, 16, 1], filter_num_list=[4, 8, 6], filter_size_list=[4, 4, 4],
                 strides_list=[2, 2, 1])
    ws = e3.keras_default_weights(small, 1)
    x = ar.make_targets(np.random.default_rng(0), 2, G=16)
    y, outs = e3.forward(small, ws, x, dtype=torch.float64, return_layers=True)
    assert [tuple(o.shape) for o in outs] == [(2, 8, 8, 8, 4), (2, 4, 4, 4, 8), (2, 4, 4, 4, 6)] and y.shape == (2, 6)
    np.testing.assert_allclose(y.numpy(), outs[-1].numpy().mean(axis=(1, 2, 3)), atol=1e-12)
    ymax = e3.forward(dict(small, final_pool='max', final_activation='sigmoid'), ws, x, dtype=torch.float64).numpy()
    np.testing.assert_allclose(ymax, 1 / (1 + np.exp(-outs[-1].numpy().max(axis=(1, 2, 3)))), atol=1e-12)
    # first layer against the definition, through BN (identity statistics at init) and ELU
    pre = e3.conv3d_same_definition(x.astype(np.float64), ws[0].astype(np.float64), 2) / np.sqrt(1 + 1e-3)
    np.testing.assert_allclose(outs[0].numpy(), np.where(pre > 0, pre, np.expm1(pre)), atol=1e-10)


def test_oracle_reproduces_committed_fixture():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with np.load(os.path.join(root, 'tests', 'golden', 'golden_enc3d_v1.npz')) as f:
        g = {k: f[k] for k in f.files}
    st = e3.MODELNET_ENCODER
    ws = e3.trained_like_weights(st, 401)
    np.testing.assert_allclose([np.asarray(w, np.float64).sum() for w in ws], g['wsum'], rtol=1e-9, atol=1e-9)
    x = ar.make_targets(np.random.Generator(np.random.PCG64(402)), 3)
    assert np.array_equal(x.reshape(3, -1).sum(1).astype(np.int64), g['x_occupancy'])
    y, outs = e3.forward(st, ws, x, return_layers=True)
    np.testing.assert_allclose(y.numpy(), g['out'], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose([float(o.double().abs().sum()) for o in outs], g['layer_abs'], rtol=1e-4)
