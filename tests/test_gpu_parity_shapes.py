"""GPU parity tests (-m gpu) for the reference's own call shapes and the decoder variants its graph admits:

* K = 32 samples of one object per decoder call (nolbo_test.py:167-177, ``sampling_num = 32``) and batch 72
  (test_modelnet_VAE_dr.py:52,124);
* a handle with max_chunk = 4096, the shape bench.py runs (256 objects x K = 16 in one chunk);
* un-rounded fp32 kernels, i.e. what ``load_weights`` of a trained reference checkpoint delivers: the GPU rounds them to
  16 bit itself, the oracle computes with the fp32 originals;
* ``activation`` in {relu, lrelu} and ``final_activation = 'None'`` (autoencoder3D.py:48-53,134-138);
* the two sigmoid forms of the fused tail (tanh.approx on the counts-only path, exp form whenever a grid or the loss is
  emitted) give the same counts up to voxels that sit on the threshold.

Tolerances as in test_gpu_parity.py (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import anytime_ref as ar, decoder_ref as dr

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2
FLIP_TOL = 1e-3


@pytest.fixture(scope='module')
def a3d_mod():
    import a3d
    return a3d


@pytest.fixture(scope='module')
def trained():
    return dr.trained_like_weights(dr.MODELNET_DECODER, 102)


def _check(mp, cnt, ref_mp, ref_cnt):
    assert np.abs(mp - ref_mp).max() < PROB_TOL
    nflip = int(((mp >= 0.5) != (ref_mp >= 0.5)).sum())
    assert nflip / mp.size < FLIP_TOL
    assert np.abs(cnt - ref_cnt).sum() <= 2 * nflip
    return nflip


def test_k32_samples_of_one_object(a3d_mod, trained):
    """nolbo_test.getPred: 32 latent draws of ONE object -> one decoder call of 32 -> mean of the 32 grids."""
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)
    dec.set_weights(trained)
    rng = np.random.default_rng(320)
    mean = dr.round_bf16(rng.standard_normal((1, 64)).astype(np.float32))
    logvar = np.full((1, 64), -2.0, np.float32)
    mu_t = np.repeat(mean, 32, 0)                                   # tf.stack([mean] * 32) nolbo_test.py:172-173
    z = a3d_mod.sampling(mu_t, np.repeat(logvar, 32, 0), seed=11, decoder=dec)
    tgt = ar.make_targets(rng, 1)
    r = a3d_mod.anytime_eval(dec, None, None, None, tgt, z_completed=z[None], return_grid=True)
    ref_mp, ref_cnt = ar.anytime_eval(dr.MODELNET_DECODER, trained, z[None], tgt)
    _check(r['mean_prob'].cpu().numpy(), r['counts'].cpu().numpy(), ref_mp, ref_cnt)
    # the plain decoder call of the same 32 latents, averaged on the host like tf.reduce_mean(axis=0) :176
    grids = dec(z)
    assert np.abs(grids.mean(0, dtype=np.float64) - ref_mp[0]).max() < PROB_TOL


def test_getPredShapes_all_instances_in_one_call(a3d_mod, trained):
    """a3d.getPredShapes = the 3-D shape block of nolbo_test.getPred (:167-182) for all selected instances at once: the
    same draws as `sampling` row by row, the mean of the 32 grids of every instance against the oracle, the
    is_sampling=False branch (decoder(inst_mean)), the empty selection, numpy in -> numpy out."""
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=64)       # 3 instances x 32 draws = 96 decodes: two chunks
    dec.set_weights(trained)
    rng = np.random.default_rng(321)
    mean = dr.round_bf16(rng.standard_normal((3, 64)).astype(np.float32))
    logvar = rng.uniform(-3.0, -1.0, (3, 64)).astype(np.float32)
    got = a3d_mod.getPredShapes(dec, mean, logvar, is_sampling=True, sampling_num=32, seed=5)
    assert isinstance(got, np.ndarray) and got.shape == (3, 64, 64, 64) and got.dtype == np.float32
    z = a3d_mod.sampling(np.repeat(mean, 32, 0), np.repeat(logvar, 32, 0), seed=5, decoder=dec).reshape(3, 32, 64)
    eps = ar.philox_normals(5, np.arange(96, dtype=np.uint64), 1, 64)[:, 0]      # row r draws with the k = 0 stream of object r
    z_ref = ar.sampling(np.repeat(mean, 32, 0), np.repeat(logvar, 32, 0), eps)
    assert np.abs(z.reshape(96, 64) - z_ref).max() < 2e-5
    ref_mp, _ = ar.anytime_eval(dr.MODELNET_DECODER, trained, z, ar.make_targets(rng, 3))
    ref_mp = ref_mp.reshape(3, 64, 64, 64)
    assert np.abs(got - ref_mp).max() < PROB_TOL
    assert ((got >= 0.5) != (ref_mp >= 0.5)).mean() < FLIP_TOL
    plain = a3d_mod.getPredShapes(dec, mean, logvar, is_sampling=False)
    ref_plain = dr.decoder_forward(dr.MODELNET_DECODER, trained, mean).numpy().reshape(3, 64, 64, 64)
    assert np.abs(plain - ref_plain).max() < PROB_TOL
    assert a3d_mod.getPredShapes(dec, np.zeros((0, 64), np.float32), np.zeros((0, 64), np.float32)).shape == (0, 64, 64, 64)


def test_anytime_eval_k32_imputed(a3d_mod, trained):
    """The composed path at the reference's sampling_num: 2 objects x K = 32 prior-sample imputations."""
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=64)
    dec.set_weights(trained)
    rng = np.random.default_rng(321)
    B, K = 2, 32
    z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
    mask = ar.bernoulli_mask(rng, B, 64, 0.5)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    tgt = ar.make_targets(rng, B)
    r = a3d_mod.anytime_eval(dec, z, mask, mu, tgt, K=K, seed=5, return_grid=True)
    zc = r['z_completed'].cpu().numpy()
    zc_ref, _ = ar.impute(z, mask, mu, K, seed=5, fill='prior_sample')
    assert np.abs(zc - zc_ref).max() < 5e-6
    ref_mp, ref_cnt = ar.anytime_eval(dr.MODELNET_DECODER, trained, zc, tgt)
    _check(r['mean_prob'].cpu().numpy(), r['counts'].cpu().numpy(), ref_mp, ref_cnt)
    c = a3d_mod.anytime_eval(dec, z, mask, mu, ar.pack_bits(tgt), K=K, seed=5)['counts']     # counts-only (tanh form)
    near = int((np.abs(r['mean_prob'].cpu().numpy() - 0.5) < 3e-4).sum())
    assert (c - r['counts']).abs().sum().item() <= 2 * near


def test_batch_72_decode(a3d_mod, trained):
    """test_modelnet_VAE_dr.py:52: batch 72 through decoder(z); 4 of the 72 checked against the oracle."""
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=96)
    dec.set_weights(trained)
    z = dr.round_bf16(np.random.default_rng(72).standard_normal((72, 64)).astype(np.float32))
    out = dec(z)
    assert out.shape == (72, 64, 64, 64, 1)
    pick = [0, 17, 64, 71]
    ref = dr.decoder_forward(dr.MODELNET_DECODER, trained, z[pick]).numpy()
    assert np.abs(out[pick] - ref).max() < PROB_TOL
    assert ((out[pick] >= 0.5) != (ref >= 0.5)).mean() < FLIP_TOL


def test_bench_shape_max_chunk_4096(a3d_mod):
    """bench.py's own shape: one handle with max_chunk = 4096, 256 objects x K = 16 in ONE chunk, Keras-default weights
    (every probability within 3e-4 of 0.5: the flip test is the binding one) and trained-like weights; objects from the
    first, a middle and the last CTA-pair item checked against the oracle."""
    rng = np.random.default_rng(4096)
    B, K = 256, 16
    z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
    mask = ar.bernoulli_mask(rng, B, 64, 0.5)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    tgt8 = ar.make_targets(rng, 8)
    bits = np.tile(ar.pack_bits(tgt8), (B // 8, 1))
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=4096)
    pick = [0, 131, 255]
    tg = np.unpackbits(bits[pick], axis=1, bitorder='little').reshape(len(pick), 64, 64, 64, 1).astype(np.float32)
    for gen, seed in ((dr.keras_default_weights, 1234), (dr.trained_like_weights, 102)):
        ws = gen(dr.MODELNET_DECODER, seed)
        dec.set_weights(ws)
        r = a3d_mod.anytime_eval(dec, z, mask, mu, bits, K=K, seed=1000, return_grid=True)
        zc = r['z_completed'][pick].cpu().numpy()
        ref_mp, ref_cnt = ar.anytime_eval(dr.MODELNET_DECODER, ws, zc, tg)
        _check(r['mean_prob'][pick].cpu().numpy(), r['counts'][pick].cpu().numpy(), ref_mp, ref_cnt)
        c = a3d_mod.anytime_eval(dec, z, mask, mu, bits, K=K, seed=1000)['counts']
        # the counts-only call differs from the grid call only on voxels whose K-mean sits on the threshold
        near = int(((r['mean_prob'] - 0.5).abs() < 3e-4).sum().item())
        assert (c - r['counts']).abs().sum().item() <= 2 * near
    dec.close()


@pytest.mark.parametrize('kind', ['trained', 'default'])
def test_unrounded_fp32_kernels(a3d_mod, kind):
    """Full fp32 kernels (bf16_kernels=False) to the GPU, the SAME fp32 arrays to the oracle: the 16-bit rounding of the
    weights is now part of the GPU's error.  SURVEY section 7 row 1 (0.19 % flips with bf16 operands at Keras-default
    init) is the case fp16 operands were chosen for."""
    if kind == 'trained':
        ws = dr.trained_like_weights(dr.MODELNET_DECODER, 104, bf16_kernels=False)
    else:
        ws = dr.keras_default_weights(dr.MODELNET_DECODER, 105, bf16_kernels=False)
    assert any(not np.array_equal(dr.round_bf16(w), w) for w in ws if w.ndim > 1)
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)
    dec.set_weights(ws)
    z = np.random.default_rng(9).standard_normal((4, 64)).astype(np.float32)      # latents un-rounded too
    out = dec(z)
    ref = dr.decoder_forward(dr.MODELNET_DECODER, ws, z).numpy()
    assert np.abs(out - ref).max() < PROB_TOL
    assert ((out >= 0.5) != (ref >= 0.5)).mean() < FLIP_TOL


@pytest.mark.parametrize('over', [{'activation': 'relu'}, {'activation': 'lrelu'}, {'final_activation': 'None'},
                                  {'activation': 'relu', 'final_activation': 'None'}])
def test_decoder_variants(a3d_mod, trained, over):
    """autoencoder3D.py:48-53 (ReLU / LeakyReLU(0.3) / ELU) and :134-138 (sigmoid or linear output)."""
    st = dict(dr.MODELNET_DECODER, **over)
    dec = a3d_mod.decoder3D(st, max_chunk=32)
    dec.set_weights(trained)
    z = dr.round_bf16(np.random.default_rng(13).standard_normal((3, 64)).astype(np.float32))
    ref, layers = dr.decoder_forward(st, trained, z, return_layers=True)
    out = dec(z)
    for li in range(5):
        g, r = dec.debug_layer(li, 3), layers[li].numpy()
        rel = np.sqrt(((g - r) ** 2).mean()) / max(np.sqrt((r ** 2).mean()), 1e-12)
        assert rel < 2e-3, f'layer {li}: relative RMS error {rel}'
    ref = ref.numpy()
    if st['final_activation'] == 'sigmoid':
        assert np.abs(out - ref).max() < PROB_TOL and ((out >= 0.5) != (ref >= 0.5)).mean() < FLIP_TOL
    else:   # logits: same relative bar as the hidden layers, and the sign (= the 0.5 threshold) agrees
        assert np.sqrt(((out - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()) < 2e-3
        assert ((out >= 0) != (ref >= 0)).mean() < FLIP_TOL


def test_sampling_and_nearest_prior_entry_points(a3d_mod):
    """a3d_sampling (function.py:35-38) and a3d_nearest_prior (nolbo.py:1488-1494) against the oracle."""
    rng = np.random.default_rng(77)
    B, D, C = 37, 64, 40
    mean = rng.standard_normal((B, D)).astype(np.float32)
    logvar = np.clip(rng.standard_normal((B, D)).astype(np.float32), -10, 10)
    z = a3d_mod.sampling(mean, logvar, seed=123, obj_offset=5)
    eps = ar.philox_normals(123, np.arange(B, dtype=np.uint64) + 5, 1, D)[:, 0]
    assert np.abs(z - ar.sampling(mean, logvar, eps)).max() < 2e-5
    mu = rng.standard_normal((C, D)).astype(np.float32)
    zc = (mu[rng.integers(0, C, B)] + 0.7 * rng.standard_normal((B, D))).astype(np.float32)
    cat = np.eye(C, dtype=np.float32)[rng.integers(0, C, B)]
    idx, acc = a3d_mod.nearest_prior(zc, mu, cat)
    ref_idx, ref_acc = ar.nearest_prior(zc, mu, cat)
    assert np.array_equal(idx.cpu().numpy(), ref_idx) and acc == pytest.approx(ref_acc, abs=1e-12)
    idx3, _ = a3d_mod.nearest_prior(np.repeat(zc[:, None, :], 3, 1) + np.arange(3, dtype=np.float32)[None, :, None], mu)
    assert np.array_equal(idx3.cpu().numpy(), ref_idx)                 # [B, K, D]: the first sample is classified
    # tie -> first minimum, like tf.argmin
    idx_t, _ = a3d_mod.nearest_prior(np.zeros((1, D), np.float32), np.ones((C, D), np.float32))
    assert int(idx_t[0]) == 0


def test_fill_none_and_geteval_without_missing_dims(a3d_mod, trained):
    """getEval's missing_prob == 0 branch applies no fill (nolbo.py:1485-1486): an exact zero in z stays zero."""
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)
    dec.set_weights(trained)
    rng = np.random.default_rng(5)
    z = rng.standard_normal((3, 64)).astype(np.float32)
    z[1, 7] = 0.0
    ones = np.ones_like(z)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    zo, _ = a3d_mod.impute(dec, z, ones, mu, K=2, seed=1, fill='none')
    assert np.array_equal(zo.cpu().numpy(), np.repeat(z[:, None, :], 2, 1))
    zm, _ = a3d_mod.impute(dec, z, ones, mu, K=1, seed=1, fill='mean')
    assert zm[1, 0, 7].item() == pytest.approx(float(mu[:, 7].mean()), rel=1e-6)      # the 'mean' fill does replace it
    tgt = ar.make_targets(rng, 3)
    cat = np.eye(40, dtype=np.float32)[rng.integers(0, 40, 3)]
    out = a3d_mod.getEval(dec, (z, tgt, cat), mu, missing_prob=0.0)
    assert np.array_equal(out[0].cpu().numpy(), dec(z))
    assert out[4] == pytest.approx(ar.nearest_prior(z, mu, cat)[1], abs=1e-12)


@pytest.mark.parametrize('n', [1, 5, 72])
def test_decode_host_formats(a3d_mod, trained, n):
    """a3d_decode_host: numpy in -> grid out in the three return formats, pageable and pinned outputs, equal to the
    device-buffer call bit for bit (fp32), to its fp16 rounding, and to its thresholding."""
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)
    dec.set_weights(trained)
    z = np.random.default_rng(n).standard_normal((n, 64)).astype(np.float32)
    ref = dec(torch.from_numpy(z).cuda()).cpu().numpy()
    a = dec(z)
    assert a.dtype == np.float32 and a.shape == (n, 64, 64, 64, 1) and np.array_equal(a, ref)
    pin = a3d_mod.pinned_empty((n, 64, 64, 64, 1), np.float32)
    b = dec(z, out=pin)
    assert np.array_equal(b, ref) and np.shares_memory(b, pin)
    h = dec(z, out_dtype='f16')
    assert h.dtype == np.float16 and np.array_equal(h, ref.astype(np.float16))
    bits = dec(z, out_dtype='bits', threshold=0.5)
    assert bits.shape == (n, 32768) and np.array_equal(bits, ar.pack_bits(ref >= 0.5))
    with pytest.raises(ValueError):
        dec(z, out=np.empty((n, 5), np.float32))
