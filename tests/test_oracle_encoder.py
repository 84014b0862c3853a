"""CPU tests of the image-encoder oracle (oracle/encoder2d_ref.py, test infrastructure) against the definition of a
Keras Conv2D('same'), the structure of the reference's Darknet19 / head2D, and the committed fixture."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder2d_ref as er

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def golden_enc():
    with np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_enc_v1.npz')) as f:
        return {k: f[k] for k in f.files}


def test_darknet19_structure_matches_reference():
    L = er.layer_list()
    convs = [l for l in L if l['kind'] == 'conv']
    assert len(convs) == 19 and sum(l['kind'] == 'maxpool' for l in L) == 5      # darknet.py:96-133 + head conv
    assert [l['filters'] for l in convs[:18]] == [32, 64, 128, 64, 128, 256, 128, 256, 512, 256, 512, 256, 512,
                                                  1024, 512, 1024, 512, 1024]
    assert [l['ksize'] for l in convs[:18]] == [3, 3, 3, 1, 3, 3, 1, 3, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3]
    assert convs[18] == {'kind': 'conv', 'filters': 32, 'ksize': 1, 'bn': False, 'act': None}   # darknet.py:155-157
    assert L[-1]['kind'] == 'global_max'                                                          # nolbo.py:783
    shapes = er.weight_shapes(L, 3)
    assert len(shapes) == 18 * 5 + 1
    assert shapes[0] == ('conv0/kernel', (3, 3, 3, 32)) and shapes[-1][1] == (1, 1, 1024, 32)
    n_params = sum(int(np.prod(s)) for n, s in shapes if n.endswith('kernel'))
    assert n_params == 19_835_744
    assert n_params == sum(k * k * ci * co for (co, k), ci in zip(
        er.DARKNET19_CONVS + [(32, 1)], [3] + [f for f, _ in er.DARKNET19_CONVS]))
    alg, dense = er.encoder_macs(L, 256, 256, 3)
    assert dense == 3_581_935_616 and alg == 3_322_454_400      # 7.16 GFLOP / image (dense)


@pytest.mark.parametrize('k', [1, 3])
def test_conv2d_same_matches_definition(k):
    rng = np.random.default_rng(k)
    x = rng.standard_normal((2, 5, 6, 3))
    w = rng.standard_normal((k, k, 3, 4))
    layers = [{'kind': 'conv', 'filters': 4, 'ksize': k, 'bn': False, 'act': None}]
    got = er.forward(layers, [w], x, dtype=torch.float64).numpy()
    np.testing.assert_allclose(got, er.conv2d_same_definition(x, w), atol=1e-12)


def test_bn_act_pool_semantics():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 4, 4, 2))
    w = rng.standard_normal((1, 1, 2, 3))
    g, b, m, v = rng.uniform(0.5, 1.5, 3), rng.standard_normal(3), rng.standard_normal(3), rng.uniform(0.5, 2, 3)
    for act, fn in (('elu', lambda t: np.where(t > 0, t, np.expm1(t))), ('relu', lambda t: np.maximum(t, 0)),
                    ('lrelu', lambda t: np.where(t > 0, t, 0.1 * t))):                 # darknet.py:87-92
        layers = [{'kind': 'conv', 'filters': 3, 'ksize': 1, 'bn': True, 'act': act}, {'kind': 'maxpool'}]
        got = er.forward(layers, [w, g, b, m, v], x, dtype=torch.float64).numpy()
        pre = x @ w[0, 0]
        y = fn(g * (pre - m) / np.sqrt(v + 1e-3) + b)                                  # Keras BN, eps = 1e-3
        want = y.reshape(1, 2, 2, 2, 2, 3).max(axis=(2, 4))                            # MaxPool2D(2, 2)
        np.testing.assert_allclose(got, want, atol=1e-12)
    layers = [{'kind': 'conv', 'filters': 3, 'ksize': 1, 'bn': False, 'act': None}, {'kind': 'global_avg'}]
    np.testing.assert_allclose(er.forward(layers, [w], x, dtype=torch.float64).numpy(), (x @ w[0, 0]).mean((1, 2)),
                               atol=1e-12)


def test_split_latent_and_sampler(golden_enc):
    e = np.array([[1., 2., 30., -40.]], np.float32)
    mean, logvar = er.split_latent(e, 2)
    assert mean.tolist() == [[1., 2.]] and logvar.tolist() == [[10., -10.]]              # nolbo.py:873
    n = er.latent_normals(5, np.arange(4096, dtype=np.uint64), 16)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1) < 0.02
    np.testing.assert_allclose(er.latent_normals(99, np.array([0, 7, 2 ** 33 + 1], np.uint64), 16),
                               golden_enc['latent_normals'], atol=1e-12)
    # the latent stream never collides with the imputation sampler's streams (k < K) of the same seed
    from oracle import anytime_ref as ar
    imp = ar.philox_normals(99, np.array([0, 7], np.uint64), 4, 16)
    assert not np.isclose(imp[:, 0], golden_enc['latent_normals'][:2]).all()


@pytest.mark.parametrize('tag,size,wseed', [('s64', 64, 301), ('s256', 256, 302)])
def test_oracle_reproduces_committed_fixture(golden_enc, tag, size, wseed):
    layers = er.layer_list()
    ws = er.trained_like_weights(layers, 3, seed=wseed, hw=size)
    np.testing.assert_allclose([np.asarray(w, np.float64).sum() for w in ws], golden_enc[f'{tag}_wsum'], rtol=1e-9,
                               atol=1e-9)
    x = np.random.Generator(np.random.PCG64(9000 + size)).uniform(0, 1, (2, size, size, 3)).astype(np.float32)
    y, outs = er.forward(layers, ws, x, return_layers=True)
    np.testing.assert_allclose(y.numpy(), golden_enc[f'{tag}_out'], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose([float(o.double().abs().sum()) for o in outs], golden_enc[f'{tag}_layer_abs'], rtol=1e-4)
