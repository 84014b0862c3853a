"""GPU parity tests (-m gpu) of the voxel encoder (encoder3D): the sm_100a CUDA path
through the C ABI against the torch-CPU oracle on identical weights and occupancy grids.  Tolerances as for the image
encoder: hidden layers within 1e-2 of the layer maximum (fp16 operands; bf16 8x), pooled output within 2e-2 absolute
on O(1) outputs."""
import numpy as np
import pytest
import torch

from oracle import anytime_ref as ar, decoder_ref as dr, encoder3d_ref as e3

pytestmark = pytest.mark.gpu

REL_TOL = 1e-2
OUT_TOL = 2e-2


@pytest.fixture(scope='module')
def a3d_mod():
    import a3d
    return a3d


@pytest.mark.parametrize('dtype,mult', [('fp16', 1), ('bf16', 8)])
def test_encoder3d_per_layer_vs_oracle(a3d_mod, dtype, mult):
    st = e3.MODELNET_ENCODER
    ws = e3.trained_like_weights(st, 41)
    x = ar.make_targets(np.random.default_rng(3), 5)                    # ragged: 5 objects, max_batch 4 -> two chunks
    ref, layers = e3.forward(st, ws, x, return_layers=True)
    enc = a3d_mod.encoder3D(st, max_batch=4, operand_dtype=dtype)
    enc.set_weights(ws)
    assert enc.output_shape == (None, 128)
    out = enc(x)
    assert out.shape == (5, 128) and out.dtype == np.float32
    for li in range(4):                                                 # the last chunk holds object 4
        got = enc.debug_layer(li, 1)
        want = layers[li].numpy()[4:5]
        rel = np.abs(got - want).max() / np.abs(want).max()
        assert rel < REL_TOL * mult, f'layer {li}: rel err {rel:.3e}'
    assert np.abs(out - ref.numpy()).max() < OUT_TOL * mult
    assert all(np.array_equal(a, b) for a, b in zip(enc.get_weights(), ws))
    enc.close()


def test_encoder3d_variants_and_chunking(a3d_mod):
    """Keras-default init, max pooling + sigmoid, no pooling, relu / lrelu activations, wider latent; chunked and
    unchunked batches agree bit-for-bit."""
    x = ar.make_targets(np.random.default_rng(5), 6)
    for st in (dict(e3.MODELNET_ENCODER, final_pool='max', final_activation='sigmoid', activation='relu'),
               dict(e3.MODELNET_ENCODER, final_pool='None', activation='lrelu', filter_num_list=[64, 128, 256, 512, 32]),
               dict(e3.MODELNET_ENCODER, filter_num_list=[64, 128, 128, 256, 400])):      # autoencoder3D.py:5-14 example
        ws = e3.trained_like_weights(st, 43)
        ref = e3.forward(st, ws, x).numpy()
        outs = []
        for mb in (4, 8):
            enc = a3d_mod.encoder3D(st, max_batch=mb)
            enc.set_weights(ws)
            outs.append(enc(torch.from_numpy(x).cuda()).cpu().numpy())
            enc.close()
        assert outs[0].shape == ref.shape
        assert np.array_equal(outs[0], outs[1])
        assert np.abs(outs[0] - ref).max() < OUT_TOL * max(np.abs(ref).max(), 1.0), st
    st = e3.MODELNET_ENCODER
    ws = e3.keras_default_weights(st, 7)
    enc = a3d_mod.encoder3D(st, max_batch=8)
    enc.set_weights(ws)
    ref = e3.forward(st, ws, x).numpy()
    assert np.abs(enc(x) - ref).max() < 1e-2 * max(np.abs(ref).max(), 1e-3)
    enc.close()


def test_voxels_to_voxels_end_to_end(a3d_mod):
    """test_modelnet_VAE path (BASELINE config 1 with its encoder): voxels -> encoder3D -> (mean, clipped logvar) ->
    sampling -> decoder -> occupancy, against the oracle chain fed the same seeded draws; then getEvalVoxels."""
    n, D = 4, 64
    est, dst = e3.MODELNET_ENCODER, dr.MODELNET_DECODER
    ews, dws = e3.trained_like_weights(est, 51), dr.trained_like_weights(dst, 102)
    x = ar.make_targets(np.random.default_rng(8), n)
    enc = a3d_mod.encoder3D(est, max_batch=n)
    enc.set_weights(ews)
    dec = a3d_mod.decoder3D(dst, max_chunk=32)
    dec.set_weights(dws)
    mean, logvar, z = enc.encode(x, D, seed=9)
    from oracle import encoder2d_ref as er
    rmean, rlogvar, rz = er.split_sample(e3.forward(est, ews, x).numpy(), D, seed=9)
    assert np.abs(mean.cpu().numpy() - rmean).max() < OUT_TOL and np.abs(logvar.cpu().numpy() - rlogvar).max() < OUT_TOL
    prob = dec(z).cpu().numpy().reshape(n, -1)
    ref_prob = dr.decoder_forward(dst, dws, rz).numpy().reshape(n, -1)
    flips = float(((prob >= 0.5) != (ref_prob >= 0.5)).mean())
    print(f'voxels->voxels: max prob err {np.abs(prob - ref_prob).max():.3e}, flipped {100 * flips:.4f} %')
    assert flips < 1e-3 and np.abs(prob - ref_prob).max() < 5e-2
    cat = np.eye(40, dtype=np.float32)[np.arange(n) % 40]
    mu = np.random.default_rng(1).standard_normal((40, D)).astype(np.float32)
    res = a3d_mod.getEvalVoxels(enc, dec, (x, x, cat), mu, missing_prob=0.5, K=2, seed=5)
    assert len(res) == 10 and res[0].shape == (n, 64, 64, 64, 1)
    enc.close()
    dec.close()


def test_errors_and_weight_io(a3d_mod, tmp_path):
    with pytest.raises(RuntimeError, match='unsupported encoder3D structure'):
        a3d_mod.encoder3D(dict(e3.MODELNET_ENCODER, strides_list=[2, 2, 2, 1, 1]))
    with pytest.raises(KeyError):
        a3d_mod.encoder3D({'name': 'x'})
    enc = a3d_mod.encoder3D(e3.MODELNET_ENCODER, max_batch=2)
    with pytest.raises(NotImplementedError):
        enc(np.zeros((1, 64, 64, 64, 1), np.float32), training=True)
    with pytest.raises(RuntimeError, match='never set'):
        enc(np.zeros((1, 64, 64, 64, 1), np.float32))
    with pytest.raises(ValueError, match='expecting 21 weights'):
        enc.set_weights([np.zeros(1)])
    ws = e3.keras_default_weights(e3.MODELNET_ENCODER, 2)
    enc.set_weights(ws)
    prefix = str(tmp_path / 'encoder3D')
    enc.save_weights(prefix, save_format='tf')
    enc2 = a3d_mod.encoder3D(e3.MODELNET_ENCODER, max_batch=2)
    enc2.load_weights(prefix)
    x = ar.make_targets(np.random.default_rng(0), 2)
    assert np.array_equal(enc(x), enc2(x))
    enc.close(); enc2.close()


def test_empty_and_single_object_batches(a3d_mod):
    st = e3.MODELNET_ENCODER
    ws = e3.keras_default_weights(st, 4)
    enc = a3d_mod.encoder3D(st, max_batch=3)
    enc.set_weights(ws)
    out0 = enc(np.zeros((0, 64, 64, 64, 1), np.float32))
    assert out0.shape == (0, 128)
    x = ar.make_targets(np.random.default_rng(11), 1)
    out1 = enc(x)
    ref = e3.forward(st, ws, x).numpy()
    assert out1.shape == (1, 128) and np.abs(out1 - ref).max() < 1e-2 * max(np.abs(ref).max(), 1e-3)
    z = np.zeros((1, 64, 64, 64, 1), np.float32)                     # an empty grid: every layer sees only its BN shift
    assert np.abs(enc(z) - e3.forward(st, ws, z).numpy()).max() < 1e-3
    enc.close()


def test_encoder3d_matches_golden_fixture(a3d_mod):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with np.load(os.path.join(root, 'tests', 'golden', 'golden_enc3d_v1.npz')) as f:
        g = {k: f[k] for k in f.files}
    st = e3.MODELNET_ENCODER
    enc = a3d_mod.encoder3D(st, max_batch=4)
    enc.set_weights(e3.trained_like_weights(st, 401))
    x = ar.make_targets(np.random.Generator(np.random.PCG64(402)), 3)
    out = enc(x)
    assert np.abs(out - g['out']).max() < OUT_TOL
    for li in range(4):
        got = enc.debug_layer(li, 3).reshape(3, -1)
        step = max(1, got.shape[1] // 64)
        want = g['layer_samples'][li]
        assert np.abs(got[:, ::step][:, :64] - want).max() < REL_TOL * max(np.abs(want).max(), 1.0)
    enc.close()
