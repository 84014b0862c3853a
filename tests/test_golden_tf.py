"""Parity against numbers the reference framework itself produced (tests/golden/golden_tf_v1.npz, written by
tests/golden/make_golden_tf.py from the unmodified reference under TensorFlow).

TensorFlow cannot be installed in the build container or on the GPU boxes (no network), so until someone runs the
generator on a machine that has it these tests SKIP LOUDLY and DESIGN.md keeps saying "parity unpinned"."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import anytime_ref as ar, decoder_ref as dr

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'golden_tf_v1.npz')
SKIP = ('PARITY UNPINNED: tests/golden/golden_tf_v1.npz is absent (TensorFlow available here: %s); run '
        'tests/golden/make_golden_tf.py on a machine with TensorFlow + the reference checkout and commit the file'
        % (importlib.util.find_spec('tensorflow') is not None))

CASES = {'mn_default': (dr.MODELNET_DECODER, dr.keras_default_weights, {}),
         'mn_trained': (dr.MODELNET_DECODER, dr.trained_like_weights, {}),
         'pa_trained': (dr.PASCAL_DECODER, dr.trained_like_weights, {}),
         'mn_relu': (dr.MODELNET_DECODER, dr.trained_like_weights, {'activation': 'relu'}),
         'mn_lrelu': (dr.MODELNET_DECODER, dr.trained_like_weights, {'activation': 'lrelu'}),
         'mn_linear': (dr.MODELNET_DECODER, dr.trained_like_weights, {'final_activation': 'None'}),
         'mn_trained_fp32': (dr.MODELNET_DECODER, lambda st, s: dr.trained_like_weights(st, s, bf16_kernels=False), {})}


@pytest.fixture(scope='module')
def tf_golden():
    if not os.path.exists(FIXTURE):
        pytest.skip(SKIP)
    with np.load(FIXTURE) as f:
        return {k: f[k] for k in f.files}


def test_generator_exists_and_names_the_unmodified_reference():
    src = open(os.path.join(HERE, 'golden', 'make_golden_tf.py')).read()
    assert 'src/net_core/autoencoder3D.py' in src and 'src/module/function.py' in src and 'import tensorflow' in src


def test_keras_variable_order_matches_the_oracle(tf_golden):
    shapes = [tuple(int(v) for v in s.split(',')) for s in tf_golden['var_shapes']]
    assert shapes == [sh for _, sh in dr.weight_shapes(dr.MODELNET_DECODER)]
    # Glorot limits of Keras' own initialisation: max |w| just below sqrt(6 / (fan_in + fan_out)), BN defaults
    for (name, sh), lo, hi in zip(dr.weight_shapes(dr.MODELNET_DECODER), tf_golden['init_min'], tf_golden['init_max']):
        if name.endswith('/kernel'):
            rec = int(np.prod(sh[:-2])) if len(sh) > 2 else 1
            fan_in, fan_out = (sh if len(sh) == 2 else (sh[-2] * rec, sh[-1] * rec))
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            assert 0.97 * lim < hi <= lim * (1 + 1e-6) and -lim * (1 + 1e-6) <= lo < -0.97 * lim
        elif name.endswith(('/gamma', '/moving_variance')):
            assert lo == hi == 1.0
        else:
            assert lo == hi == 0.0


@pytest.mark.parametrize('tag', sorted(CASES))
def test_oracle_matches_tensorflow(tf_golden, tag):
    st0, gen, over = CASES[tag]
    st = dict(st0, **over)
    ws = gen(st0, int(tf_golden[f'{tag}_wseed']))
    out, layers = dr.decoder_forward(st, ws, tf_golden[f'{tag}_z'], return_layers=True)
    flat = out.numpy().reshape(2, -1)
    assert np.abs(flat[:, tf_golden['sample_idx']] - tf_golden[f'{tag}_prob_samples']).max() < 2e-5
    bits = np.unpackbits(tf_golden[f'{tag}_bits'], axis=1, bitorder='little').astype(bool)
    assert ((flat >= 0.5) != bits).mean() < 1e-5
    for li in range(5):
        l = layers[li].numpy().reshape(2, -1)
        got = l[:, tf_golden[f'{tag}_layer{li}_idx']]
        want = tf_golden[f'{tag}_layer{li}_samples']
        assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())
    if f'{tag}_counts' in tf_golden:
        tgt = np.unpackbits(tf_golden[f'{tag}_target_bits'], axis=1, bitorder='little').reshape(2, 64, 64, 64, 1)
        # voxelPrecisionRecall / binary_loss of function.py on TensorFlow's own output grid
        tf_bits = bits.reshape(2, 64, 64, 64, 1).astype(np.float32)
        assert np.array_equal(ar.counts(tgt.astype(np.float32), tf_bits, 0.5), tf_golden[f'{tag}_counts'].astype(np.int64))
        np.testing.assert_allclose(ar.binary_loss(out.numpy(), tgt.astype(np.float32), gamma=0.6), tf_golden[f'{tag}_bce'],
                                   rtol=1e-3)


def test_sampling_moments(tf_golden):
    m, s = tf_golden['sampling_mean_std']
    assert abs(m - 0.5) < 0.05 and abs(s - 2.0) < 0.05


def test_checkpoint_reader_on_a_file_tensorflow_wrote(tf_golden, tmp_path):
    import importlib
    tfc = importlib.import_module('anytime-3d-reconstruction_b200.tf_checkpoint')
    prefix = str(tmp_path / 'ckpt')
    open(prefix + '.index', 'wb').write(tf_golden['ckpt_index'].tobytes())
    open(prefix + '.data-00000-of-00001', 'wb').write(tf_golden['ckpt_data'].tobytes())
    got = tfc.load_keras_weights(prefix)
    assert len(got) == int(tf_golden['ckpt_n'])
    for i, w in enumerate(got):
        assert np.array_equal(w, tf_golden[f'ckpt_w{i}'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', sorted(CASES))
def test_cuda_path_matches_tensorflow(tf_golden, tag):
    import a3d
    st0, gen, over = CASES[tag]
    st = dict(st0, **over)
    dec = a3d.decoder3D(st, max_chunk=32)
    dec.set_weights(gen(st0, int(tf_golden[f'{tag}_wseed'])))
    flat = dec(tf_golden[f'{tag}_z']).reshape(2, -1)
    want = tf_golden[f'{tag}_prob_samples']
    bits = np.unpackbits(tf_golden[f'{tag}_bits'], axis=1, bitorder='little').astype(bool)
    if st['final_activation'] == 'sigmoid':
        assert np.abs(flat[:, tf_golden['sample_idx']] - want).max() < 1e-2
        assert ((flat >= 0.5) != bits).mean() < 1e-3
    else:
        assert np.abs(flat[:, tf_golden['sample_idx']] - want).max() < 4e-3 * max(1.0, np.abs(want).max())
