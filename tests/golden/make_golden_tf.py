"""Writes tests/golden/golden_tf_v1.npz from the UNMODIFIED reference running under TensorFlow.

    python tests/golden/make_golden_tf.py [--reference /root/reference] [--out tests/golden/golden_tf_v1.npz]

This is the parity anchor SURVEY.md section 8c asks for.  It cannot run in the build container or on the GPU boxes
(TensorFlow is not installed and there is no network: `importlib.util.find_spec('tensorflow')` is None, which is what
bench.py reports as `tf_available`), so the fixture is NOT committed yet and tests/test_golden_tf.py skips loudly.
Run it once on any machine that has TensorFlow 2.x (CPU is enough, ~1 min) next to a checkout of the reference and
commit the .npz: from then on the oracle (CPU suite) and the CUDA path (-m gpu suite) are checked against numbers the
reference framework itself produced.

What it records (all inputs seeded; weights come from the same seeded generators the tests use, fed to the Keras
model with set_weights, so the 45 MB weight sets are not stored):
  * the Keras variable order / shapes / names of `decoder3D(structure)` (src/net_core/autoencoder3D.py:104-139);
  * statistics of Keras' own default initialisation (pins the Glorot limits the oracle's generator restates);
  * for three weight sets x {elu, relu, lrelu, final 'None'} variants: z, sampled per-layer activations and sums, sampled
    output probabilities and the bit-packed thresholded grid of `decoder(z, training=False)`;
  * `voxelPrecisionRecall`, `binary_loss(gamma=0.6)` and `sampling` of src/module/function.py:35-38,73-82,100-115 on
    those outputs and seeded targets;
  * a small Keras model's `save_weights(prefix)` TF-format checkpoint (index + data bytes) with its `get_weights()`,
    which pins the checkpoint reader of anytime-3d-reconstruction_b200/tf_checkpoint.py against a file TensorFlow wrote.
"""
import argparse
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def load_reference_module(ref_root: str, rel: str, name: str):
    """Import one file of the reference by path, unmodified."""
    path = os.path.join(ref_root, rel)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default=os.environ.get('A3D_REFERENCE', '/root/reference'))
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden', 'golden_tf_v1.npz'))
    args = ap.parse_args()
    if importlib.util.find_spec('tensorflow') is None:
        sys.exit('TensorFlow is not installed: this generator needs the reference framework itself (see the docstring)')
    import tensorflow as tf
    if importlib.util.find_spec('cv2') is None:       # function.py imports cv2 at module level and never uses it on this path
        sys.modules['cv2'] = types.ModuleType('cv2')
    ae = load_reference_module(args.reference, 'src/net_core/autoencoder3D.py', 'ref_autoencoder3D')
    fn = load_reference_module(args.reference, 'src/module/function.py', 'ref_function')
    from oracle import anytime_ref as ar, decoder_ref as dr

    out = {'tf_version': np.array(tf.__version__), 'keras_version': np.array(getattr(tf.keras, '__version__', '?'))}
    rng = np.random.Generator(np.random.PCG64(20261018))
    sample_idx = np.sort(rng.choice(64 ** 3, 4096, replace=False)).astype(np.int64)
    out['sample_idx'] = sample_idx

    def build(structure):
        tf.keras.backend.clear_session()
        return ae.decoder3D(structure)

    # ---- variable order, shapes, names; Keras' own default initialisation
    tf.keras.utils.set_random_seed(1)
    model = build(dict(dr.MODELNET_DECODER))
    out['var_names'] = np.array([v.name for v in model.weights])
    out['var_shapes'] = np.array([','.join(str(int(d)) for d in v.shape) for v in model.weights])
    init = model.get_weights()
    out['init_min'] = np.array([float(w.min()) for w in init])
    out['init_max'] = np.array([float(w.max()) for w in init])
    out['init_std'] = np.array([float(w.std()) for w in init])

    # ---- decoder outputs on the shared seeded weight sets
    cases = [('mn_default', dr.MODELNET_DECODER, 101, dr.keras_default_weights, {}),
             ('mn_trained', dr.MODELNET_DECODER, 102, dr.trained_like_weights, {}),
             ('pa_trained', dr.PASCAL_DECODER, 103, dr.trained_like_weights, {}),
             ('mn_relu', dr.MODELNET_DECODER, 102, dr.trained_like_weights, {'activation': 'relu'}),
             ('mn_lrelu', dr.MODELNET_DECODER, 102, dr.trained_like_weights, {'activation': 'lrelu'}),
             ('mn_linear', dr.MODELNET_DECODER, 102, dr.trained_like_weights, {'final_activation': 'None'}),
             # un-rounded fp32 kernels (what a trained checkpoint holds): the oracle's generators round to bf16 by default
             ('mn_trained_fp32', dr.MODELNET_DECODER, 104, lambda st, s: dr.trained_like_weights(st, s, bf16_kernels=False), {})]
    for tag, st0, wseed, gen, over in cases:
        st = dict(st0, **over)
        ws = gen(st0, wseed)
        model = build(st)
        model.set_weights(ws)
        D = st['input_dim']
        z = dr.round_bf16(rng.standard_normal((2, D)).astype(np.float32))
        # per-layer activations: the outputs of the activation layers (ELU / ReLU / LeakyReLU) in graph order
        act_types = (tf.keras.layers.ELU, tf.keras.layers.ReLU, tf.keras.layers.LeakyReLU)
        taps = [l.output for l in model.layers if isinstance(l, act_types)]
        probe = tf.keras.Model(model.inputs, taps + [model.output])
        vals = [np.asarray(v) for v in probe(z, training=False)]
        prob = vals[-1].reshape(2, -1)
        out[f'{tag}_wseed'] = np.int64(wseed)
        out[f'{tag}_z'] = z
        out[f'{tag}_prob_samples'] = prob[:, sample_idx].astype(np.float32)
        out[f'{tag}_bits'] = ar.pack_bits(prob >= 0.5)
        out[f'{tag}_layer_sum'] = np.array([float(np.asarray(v, np.float64).sum()) for v in vals[:-1]])
        out[f'{tag}_layer_abs'] = np.array([float(np.abs(np.asarray(v, np.float64)).sum()) for v in vals[:-1]])
        for li, v in enumerate(vals[:-1]):
            flat = v.reshape(2, -1)
            idx = sample_idx[sample_idx < flat.shape[1]][:1024]
            out[f'{tag}_layer{li}_idx'] = idx
            out[f'{tag}_layer{li}_samples'] = flat[:, idx].astype(np.float32)
        if tag in ('mn_trained', 'mn_default'):
            tgt = ar.make_targets(rng, 2)
            tp, fp, fnn = fn.voxelPrecisionRecall(tf.constant(tgt), tf.constant(vals[-1]), 0.5)
            out[f'{tag}_target_bits'] = ar.pack_bits(tgt)
            out[f'{tag}_counts'] = np.stack([np.asarray(tp), np.asarray(fp), np.asarray(fnn)], -1).astype(np.float64)
            out[f'{tag}_bce'] = np.asarray(fn.binary_loss(tf.constant(vals[-1]), tf.constant(tgt), gamma=0.60),
                                           np.float64)

    # ---- sampling(): moments only (the reference draws from an unseeded tf.random.normal)
    mu = np.full((4096, 16), 0.5, np.float32)
    lv = np.full((4096, 16), np.log(4.0), np.float32)
    s = np.asarray(fn.sampling(tf.constant(mu), tf.constant(lv)))
    out['sampling_mean_std'] = np.array([s.mean(), s.std()], np.float64)       # expect 0.5, 2.0

    # ---- a TF-format checkpoint written by Keras itself (nolbo.py:1572-1574 uses save_weights(prefix))
    tf.keras.utils.set_random_seed(2)
    small = tf.keras.Sequential([tf.keras.layers.Input(shape=[5]), tf.keras.layers.Dense(7),
                                 tf.keras.layers.BatchNormalization(), tf.keras.layers.Dense(3, use_bias=False)])
    small(np.zeros((1, 5), np.float32))
    with tempfile.TemporaryDirectory() as td:
        prefix = os.path.join(td, 'ckpt')
        small.save_weights(prefix)
        out['ckpt_index'] = np.frombuffer(open(prefix + '.index', 'rb').read(), np.uint8)
        out['ckpt_data'] = np.frombuffer(open(prefix + '.data-00000-of-00001', 'rb').read(), np.uint8)
    for i, w in enumerate(small.get_weights()):
        out[f'ckpt_w{i}'] = w
    out['ckpt_n'] = np.int64(len(small.get_weights()))

    np.savez_compressed(args.out, **out)
    print('wrote', args.out, os.path.getsize(args.out), 'bytes; TensorFlow', tf.__version__)


if __name__ == '__main__':
    main()
