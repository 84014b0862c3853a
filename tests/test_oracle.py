"""CPU tests of the oracle (test infrastructure) against definitions, known answers and the committed fixtures."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import anytime_ref as ar, decoder_ref as dr, numpy_ref as nr


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], '6627e8d5 e169c58d bc57ac4c 9b00dbd8'),
        ([0xffffffff] * 4, [0xffffffff] * 2, '408f276d 41c83b0e a20bc7c6 6d5451fd'),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         'd16cfe09 94fdcceb 5001e420 24126ea1'),
    ]
    for c, k, want in kat:
        r = ar.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert ' '.join('%08x' % v for v in r) == want


def test_philox_normals_statistics_and_counter_layout(golden):
    n = ar.philox_normals(5, np.arange(64, dtype=np.uint64), 8, 64)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1.0) < 0.02
    # sharding invariance: object ids, not positions, key the stream
    a = ar.philox_normals(5, np.arange(10, 20, dtype=np.uint64), 2, 16)
    b = ar.philox_normals(5, np.arange(0, 20, dtype=np.uint64), 2, 16)[10:]
    assert np.array_equal(a, b)
    assert np.array_equal(ar.philox_words(0x1234ABCD5678EF01, np.array([0, 1, 2 ** 33 + 5], np.uint64), 3, 10),
                          golden['philox_words'])
    np.testing.assert_allclose(ar.philox_normals(77, np.arange(4, dtype=np.uint64) + 1000, 2, 16),
                               golden['philox_normals'], rtol=0, atol=1e-12)


def test_structure_arithmetic_matches_reference_quirks():
    s = dr.parse_structure(dr.MODELNET_DECODER)
    assert s['grid0'] == [4, 4, 4] and s['ch0'] == 8 and s['dense_units'] == 512
    shapes = dr.weight_shapes(dr.MODELNET_DECODER)
    assert len(shapes) == 27
    assert sum(int(np.prod(sh)) for _, sh in shapes) == 11_315_456
    assert shapes[6][1] == (4, 4, 4, 512, 8) and shapes[26][1] == (4, 4, 4, 1, 64)
    assert dr.weight_shapes(dr.PASCAL_DECODER)[0][1] == (16, 512)


@pytest.mark.parametrize('stride,n_in', [(1, 4), (2, 4), (2, 5)])
def test_transposed_conv_is_adjoint_of_same_forward_conv(stride, n_in):
    """Definition check independent of both restatements: <ConvT(x), y> == <x, Conv_SAME(y)> with TF SAME padding."""
    rng = np.random.default_rng(stride * 10 + n_in)
    cin, cout, k = 3, 2, 4
    w = rng.standard_normal((k, k, k, cout, cin))
    x = rng.standard_normal((1, n_in, n_in, n_in, cin))
    big = n_in * stride
    y = rng.standard_normal((1, big, big, big, cout))
    up = dr.conv3d_transpose_same(torch.from_numpy(x).permute(0, 4, 1, 2, 3), torch.from_numpy(w), stride)
    up = up.permute(0, 2, 3, 4, 1).numpy()
    np.testing.assert_allclose(up, nr.conv3d_transpose_same(x, w, stride), atol=1e-12)
    # forward SAME conv of y (big -> small), kernel [k,k,k,in=cout,out=cin]
    out = -(-big // stride)
    pad_total = max((out - 1) * stride + k - big, 0)
    pb, pa = pad_total // 2, pad_total - pad_total // 2
    yt = F.pad(torch.from_numpy(y).permute(0, 4, 1, 2, 3), (pb, pa, pb, pa, pb, pa))
    wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2)  # [out=cin, in=cout, k,k,k]
    fwd = F.conv3d(yt, wt, stride=stride).permute(0, 2, 3, 4, 1).numpy()
    assert fwd.shape == x.shape
    np.testing.assert_allclose((up * y).sum(), (x * fwd).sum(), rtol=1e-10)


def test_torch_oracle_matches_fp64_definition_on_small_decoder():
    small = dict(dr.MODELNET_DECODER, input_dim=8, filter_num_list=[16, 8, 8, 4, 1], output_shape=[32, 32, 32, 1],
                 strides_list=[1, 2, 2, 2, 1])
    ws = dr.keras_default_weights(small, 1)
    rng = np.random.default_rng(0)
    names = [n for n, _ in dr.weight_shapes(small)]
    for i, n in enumerate(names):
        if 'bn' in n:
            ws[i] = rng.uniform(0.5, 1.5, ws[i].shape).astype(np.float32)
    z = rng.standard_normal((2, 8)).astype(np.float32)
    s = dr.parse_structure(small)
    o1, l1 = dr.decoder_forward(small, ws, z, dtype=torch.float64, return_layers=True)
    o2, l2 = nr.decoder_forward(ws, z, s['strides'], s['grid0'], s['ch0'], return_layers=True)
    for a, b in zip(l1, l2):
        np.testing.assert_allclose(a.numpy(), b, atol=1e-12)
    np.testing.assert_allclose(o1.numpy(), o2, atol=1e-12)
    o32 = dr.decoder_forward(small, ws, z)
    assert np.abs(o32.numpy() - o2).max() < 1e-5


@pytest.mark.parametrize('tag,st,gen', [('mn_default', dr.MODELNET_DECODER, dr.keras_default_weights),
                                        ('mn_trained', dr.MODELNET_DECODER, dr.trained_like_weights),
                                        ('pa_trained', dr.PASCAL_DECODER, dr.trained_like_weights)])
def test_oracle_against_golden_fixture(golden, tag, st, gen):
    ws = gen(st, int(golden[f'{tag}_wseed']))
    np.testing.assert_allclose([np.asarray(w, np.float64).sum() for w in ws], golden[f'{tag}_wsum'], rtol=1e-9, atol=1e-9)
    prob, layers = dr.decoder_forward(st, ws, golden[f'{tag}_z'], return_layers=True)
    prob = prob.numpy().reshape(2, -1)
    np.testing.assert_allclose(prob[:, golden['sample_idx']], golden[f'{tag}_prob_samples'], atol=2e-6)
    flips = ar.pack_bits(prob >= 0.5) ^ golden[f"{tag}_bits"]
    assert np.unpackbits(flips).mean() < 2e-5   # thread-count dependent fp32 summation order only
    np.testing.assert_allclose([float(l.double().abs().sum()) for l in layers], golden[f'{tag}_layer_abs'], rtol=1e-5)


def test_default_init_statistics():
    ws = dr.keras_default_weights(dr.MODELNET_DECODER, 3)
    k = ws[11]   # convT2 kernel [4,4,4,256,512]
    lim = np.sqrt(6.0 / (64 * 256 + 64 * 512))
    assert abs(np.abs(k).max() - lim) / lim < 0.01
    assert np.array_equal(dr.round_bf16(k), k)          # kernels are bf16-representable
    assert np.all(ws[2] == 1) and np.all(ws[3] == 0) and np.all(ws[4] == 0) and np.all(ws[5] == 1)


def test_impute_semantics(golden):
    z, mask, mu = golden['imp_z'], golden['imp_mask'], golden['imp_mu']
    for fill in ('prior_sample', 'mean', 'normal'):
        zo, cs = ar.impute(z, mask, mu, 3, seed=4242, obj_offset=7, fill=fill)
        np.testing.assert_allclose(zo, golden[f'imp_{fill}_z'], atol=1e-6)
        assert np.array_equal(cs, golden[f'imp_{fill}_c'])
    zo, cs = ar.impute(z, mask, mu, 3, seed=4242, obj_offset=7, fill='mean')
    pm = mu.astype(np.float64).mean(0).astype(np.float32)
    assert np.allclose(zo[1, 0], pm)                       # nothing received -> prior mean everywhere
    assert np.array_equal(zo[2, 0], z[2])                  # everything received -> untouched
    assert zo[0, 0, 3] == pm[3]                            # a received exact zero is overwritten (where(z == 0) quirk)
    zp, _ = ar.impute(z, mask, mu, 3, seed=4242, obj_offset=7, fill='prior_sample')
    keep = mask[:, None, :] == 1
    assert np.array_equal(np.broadcast_to(zo, zp.shape)[np.broadcast_to(keep, zp.shape)],
                          zp[np.broadcast_to(keep, zp.shape)])
    assert zp[1].std() > 0.5                               # missing dims get N(mu_c*, 1) draws, different per sample
    assert not np.array_equal(zp[1, 0], zp[1, 1])
    # full mask == missing_prob 0 branch (nolbo.py:1485-1486)
    assert np.all(ar.bernoulli_mask(np.random.default_rng(0), 3, 4, 0.0) == 1)
    assert np.array_equal(ar.prefix_mask(2, 5, [0, 3]), [[0, 0, 0, 0, 0], [1, 1, 1, 0, 0]])


def test_counts_edge_cases():
    t = np.zeros((4, 64), np.float32)
    p = np.zeros((4, 64), np.float32)
    t[1] = 1
    p[2] = 1
    t[3, :10] = 1
    p[3, 5:20] = 0.5     # ties count as occupied (>=)
    c = ar.counts(t, p, 0.5)
    assert c.tolist() == [[0, 0, 0], [0, 0, 64], [0, 64, 0], [5, 10, 5]]
    tp, fp, fn = nr.voxel_precision_recall(t, p, 0.5)
    assert np.array_equal(np.stack([tp, fp, fn], -1), c)
    assert ar.iou_from_counts(c)[1] == pytest.approx(5 / (5 + 74 + 69))
    bits = ar.pack_bits(t)
    assert bits.shape == (4, 8) and bits[3, 0] == 0xFF and bits[3, 1] == 0x03


def test_anytime_eval_golden(golden):
    st = dr.MODELNET_DECODER
    ws = dr.trained_like_weights(st, 102)
    tgt = np.unpackbits(golden['ev_target_bits'], axis=1, bitorder='little').reshape(2, 64, 64, 64, 1)
    mp, cnt = ar.anytime_eval(st, ws, golden['ev_zc'], tgt)
    assert np.abs(cnt - golden['ev_counts']).max() <= 3
    np.testing.assert_allclose(mp.reshape(2, -1)[:, golden['sample_idx']], golden['ev_mean_samples'], atol=2e-6)
