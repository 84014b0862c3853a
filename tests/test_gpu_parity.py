"""GPU parity tests (-m gpu): the sm_100a CUDA path, called through the C ABI (ctypes), against the CPU oracle on
identical weights, latents and masks.

Tolerances (BASELINE.json north_star): occupancy probabilities within 1e-2 max-abs, fewer than 0.1 % of thresholded
voxels differing; integer / index work (Philox words -> masks, counts, packing) bit-exact; Box-Muller normals within
5e-6 absolute (fp32 logf/sincospif vs the fp64 evaluation of the same fp32 uniforms)."""
import numpy as np
import pytest
import torch

from oracle import anytime_ref as ar, decoder_ref as dr

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2       # max |p_gpu - p_oracle|
FLIP_TOL = 1e-3       # fraction of voxels on the other side of the 0.5 threshold
NORMAL_TOL = 5e-6


@pytest.fixture(scope='module')
def a3d_mod():
    import a3d
    return a3d


@pytest.fixture(scope='module')
def weights():
    return {
        ('mn', 'default'): dr.keras_default_weights(dr.MODELNET_DECODER, 101),
        ('mn', 'trained'): dr.trained_like_weights(dr.MODELNET_DECODER, 102),
        ('pa', 'trained'): dr.trained_like_weights(dr.PASCAL_DECODER, 103),
    }


@pytest.fixture(scope='module')
def decoders(a3d_mod, weights):
    out = {}
    for (ds, kind), ws in weights.items():
        st = dr.MODELNET_DECODER if ds == 'mn' else dr.PASCAL_DECODER
        d = a3d_mod.decoder3D(st, max_chunk=64)
        d.set_weights(ws)
        out[(ds, kind)] = d
    return out


def flips(a, b, thr=0.5):
    return float(((a >= thr) != (b >= thr)).mean())


# ------------------------------------------------------------------------------------------------ decoder
@pytest.mark.parametrize('key', [('mn', 'default'), ('mn', 'trained'), ('pa', 'trained')])
def test_decoder_matches_golden_fixture(golden, decoders, key):
    tag = f'{key[0]}_{key[1]}'
    dec = decoders[key]
    out = dec(golden[f'{tag}_z'])
    assert out.shape == (2, 64, 64, 64, 1) and out.dtype == np.float32
    flat = out.reshape(2, -1)
    assert np.abs(flat[:, golden['sample_idx']] - golden[f'{tag}_prob_samples']).max() < PROB_TOL
    gbits = np.unpackbits(golden[f'{tag}_bits'], axis=1, bitorder='little')
    assert ((flat >= 0.5) != gbits.astype(bool)).mean() < FLIP_TOL


@pytest.mark.parametrize('key', [('mn', 'default'), ('mn', 'trained'), ('pa', 'trained')])
def test_decoder_per_layer_and_end_to_end_vs_oracle(decoders, weights, key):
    dec, ws = decoders[key], weights[key]
    st = dec.structure
    rng = np.random.default_rng(7)
    n = 5     # ragged: not a multiple of any tile size
    z = dr.round_bf16(rng.standard_normal((n, st['input_dim'])).astype(np.float32))
    ref, layers = dr.decoder_forward(st, ws, z, return_layers=True)
    out = dec(z)
    for li in range(5):
        g = dec.debug_layer(li, n)
        r = layers[li].numpy()
        rel = np.sqrt(((g - r) ** 2).mean()) / np.sqrt((r ** 2).mean())
        assert rel < 2e-3, f'layer {li}: relative RMS error {rel}'
        assert np.abs(g - r).max() < 4e-3 * max(1.0, np.abs(r).max()), f'layer {li}'
    ref = ref.numpy()
    assert np.abs(out - ref).max() < PROB_TOL
    assert flips(out, ref) < FLIP_TOL


def test_decoder_call_surface(a3d_mod, decoders, weights):
    dec = decoders[('mn', 'trained')]
    with pytest.raises(NotImplementedError):
        dec(np.zeros((1, 64), np.float32), training=True)
    got = dec.get_weights()
    assert len(got) == 27 and all(np.array_equal(a, b) for a, b in zip(got, weights[('mn', 'trained')]))
    with pytest.raises(ValueError):
        dec.set_weights(got[:-1])
    bad = [w.copy() for w in got]
    bad[6] = bad[6][..., :4]
    with pytest.raises(ValueError):
        dec.set_weights(bad)
    z = np.random.default_rng(0).standard_normal((3, 64)).astype(np.float32)
    a = dec(z)
    b = dec(torch.from_numpy(z).cuda())
    assert isinstance(b, torch.Tensor) and b.is_cuda and np.array_equal(a, b.cpu().numpy())
    assert dec(np.zeros((0, 64), np.float32)).shape == (0, 64, 64, 64, 1)       # empty batch
    with pytest.raises(RuntimeError, match='unsupported decoder structure'):
        a3d_mod.decoder3D(dict(dr.MODELNET_DECODER, filter_num_list=[512, 256, 128, 32, 1]))
    fresh = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)
    with pytest.raises(RuntimeError, match='never set'):
        fresh(z)


def test_decoder_batch_and_chunk_invariance(a3d_mod, weights):
    """Size-independent property: a latent decodes to bit-identical output wherever it sits in a batch / chunk."""
    ws = weights[('mn', 'trained')]
    rng = np.random.default_rng(3)
    z = rng.standard_normal((70, 64)).astype(np.float32)
    small = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)    # 70 decodes -> 3 chunks
    big = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=96)      # one chunk
    small.set_weights(ws)
    big.set_weights(ws)
    a = small(torch.from_numpy(z).cuda())
    b = big(torch.from_numpy(z).cuda())
    assert torch.equal(a, b)
    c = big(torch.from_numpy(z[::-1].copy()).cuda())
    assert torch.equal(c.flip(0), b)


def test_simt_and_tcgen05_paths_agree(a3d_mod, weights):
    ws = weights[('mn', 'trained')]
    z = np.random.default_rng(1).standard_normal((3, 64)).astype(np.float32)
    t = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32, impl='tcgen05')
    s = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32, impl='simt')
    t.set_weights(ws)
    s.set_weights(ws)
    a, b = t(z), s(z)
    assert np.abs(a - b).max() < 5e-3 and flips(a, b) < 2e-4


def test_bf16_operands_mode(a3d_mod, weights):
    ws = weights[('mn', 'trained')]
    z = dr.round_bf16(np.random.default_rng(2).standard_normal((4, 64)).astype(np.float32))
    d = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32, operand_dtype='bf16')
    d.set_weights(ws)
    out = d(z)
    ref = dr.decoder_forward(dr.MODELNET_DECODER, ws, z).numpy()
    assert flips(out, ref) < FLIP_TOL          # bf16 activations: threshold parity holds, probabilities are looser
    assert np.abs(out - ref).max() < 5e-2


# ------------------------------------------------------------------------------------------------ sampler
@pytest.mark.parametrize('fill', ['prior_sample', 'mean', 'normal'])
def test_impute_matches_oracle_and_golden(a3d_mod, decoders, golden, fill):
    dec = decoders[('pa', 'trained')]       # D = 16
    z, mask, mu = golden['imp_z'], golden['imp_mask'], golden['imp_mu']
    zo, cs = a3d_mod.impute(dec, z, mask, mu, K=3, seed=4242, obj_offset=7, fill=fill)
    zo, cs = zo.cpu().numpy(), cs.cpu().numpy()
    assert np.abs(zo - golden[f'imp_{fill}_z']).max() < NORMAL_TOL
    assert np.array_equal(cs, golden[f'imp_{fill}_c'])
    if fill != 'normal':
        keep = np.broadcast_to(mask[:, None, :] == 1, zo.shape) & np.broadcast_to(z[:, None, :] != 0, zo.shape)
        assert np.array_equal(zo[keep], np.broadcast_to(z[:, None, :], zo.shape)[keep])   # received dims untouched, bit-exact


@pytest.mark.parametrize('B,K,D', [(1, 1, 64), (3, 5, 64), (257, 2, 16), (64, 32, 64)])
def test_impute_shapes_and_sharding_invariance(a3d_mod, decoders, B, K, D):
    dec = decoders[('mn', 'trained') if D == 64 else ('pa', 'trained')]
    rng = np.random.default_rng(B * 100 + K)
    z = rng.standard_normal((B, D)).astype(np.float32)
    mask = ar.bernoulli_mask(rng, B, D, 0.6)
    mu = rng.standard_normal((12, D)).astype(np.float32)
    zo, cs = a3d_mod.impute(dec, z, mask, mu, K=K, seed=2 ** 40 + 17, obj_offset=2 ** 33, fill='prior_sample')
    ref, rcs = ar.impute(z, mask, mu, K, seed=2 ** 40 + 17, obj_offset=2 ** 33, fill='prior_sample')
    assert np.abs(zo.cpu().numpy() - ref).max() < NORMAL_TOL
    assert (cs.cpu().numpy() == rcs).mean() > 0.99          # argmin ties in fp32 vs fp64 only
    if B > 2:   # a shard with the matching object offset reproduces the slice of the global result bit-for-bit
        lo = B // 2
        part, _ = a3d_mod.impute(dec, z[lo:], mask[lo:], mu, K=K, seed=2 ** 40 + 17, obj_offset=2 ** 33 + lo)
        assert torch.equal(part, zo[lo:])


def test_sampling_reference_signature(a3d_mod):
    mu = np.full((2000, 16), 3.0, np.float32)
    lv = np.full((2000, 16), np.log(4.0), np.float32)
    s = a3d_mod.sampling(mu, lv, seed=5)
    assert s.shape == mu.shape and abs(s.mean() - 3.0) < 0.05 and abs(s.std() - 2.0) < 0.05
    assert np.array_equal(s, a3d_mod.sampling(mu, lv, seed=5))
    assert not np.array_equal(s, a3d_mod.sampling(mu, lv, seed=6))


# ------------------------------------------------------------------------------------------------ scoring
def test_voxel_precision_recall_exact(a3d_mod):
    rng = np.random.default_rng(11)
    B = 6
    t = (rng.random((B, 64, 64, 64, 1)) < 0.2).astype(np.float32)
    p = rng.random((B, 64, 64, 64, 1)).astype(np.float32)
    t[0] = 0; p[1] = 0; t[2] = 1; p[3] = 1
    p[4].reshape(-1)[:1000] = 0.5                        # exact ties are occupied (>=)
    tp, fp, fn = a3d_mod.voxelPrecisionRecall(t, p)
    ref = ar.counts(t, p, 0.5)
    assert np.array_equal(np.stack([tp, fp, fn], -1).astype(np.int64), ref)
    tp2, fp2, fn2 = a3d_mod.voxelPrecisionRecall(t, p, prob=0.3)
    assert np.array_equal(np.stack([tp2, fp2, fn2], -1).astype(np.int64), ar.counts(t, p, 0.3))
    from a3d import pack_targets
    d = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=32)
    assert np.array_equal(pack_targets(d, t).cpu().numpy(), ar.pack_bits(t))


def test_binary_loss_and_threshold_sweep(a3d_mod):
    rng = np.random.default_rng(12)
    B = 3
    t = (rng.random((B, 64, 64, 64, 1)) < 0.15).astype(np.float32)
    p = rng.random((B, 64, 64, 64, 1)).astype(np.float32)
    p[0].reshape(-1)[:500] = 0.0          # clipped to 1e-7
    p[1].reshape(-1)[:500] = 1.0          # clipped to 1 - 1e-7
    got = a3d_mod.binary_loss(p, t, gamma=0.6)
    ref = ar.binary_loss(p, t, gamma=0.6)
    np.testing.assert_allclose(got, ref, rtol=2e-6)
    with pytest.raises(NotImplementedError):
        a3d_mod.binary_loss(p, t, b_range=True)
    thr = [(i + 1) / 20 for i in range(19)]                     # notebook: r2 with div = 20
    c = a3d_mod.voxelPrecisionRecallSweep(t, p, thr)
    assert np.array_equal(c, ar.counts_sweep(t, p, thr, strict=True))
    c2 = a3d_mod.voxelPrecisionRecallSweep(t, p, [0.5], strict=False)
    assert np.array_equal(c2[:, 0], ar.counts(t, p, 0.5))


# ------------------------------------------------------------------------------------------------ fused anytime path
def test_anytime_eval_golden(a3d_mod, decoders, golden):
    dec = decoders[('mn', 'trained')]
    bits = golden['ev_target_bits']
    r = a3d_mod.anytime_eval(dec, golden['ev_z'], golden['ev_mask'], golden['ev_mu'], bits, K=2, seed=9,
                             return_grid=True)
    assert np.abs(r['z_completed'].cpu().numpy() - golden['ev_zc']).max() < NORMAL_TOL
    mp = r['mean_prob'].cpu().numpy().reshape(2, -1)
    assert np.abs(mp[:, golden['sample_idx']] - golden['ev_mean_samples']).max() < PROB_TOL
    gb = np.unpackbits(golden['ev_mean_bits'], axis=1, bitorder='little').astype(bool)
    nflip = int(((mp >= 0.5) != gb).sum())
    assert nflip / gb.size < FLIP_TOL
    assert np.abs(r['counts'].cpu().numpy() - golden['ev_counts']).sum() <= 2 * nflip + 2


@pytest.mark.parametrize('B,K,chunk,fill', [(3, 4, 32, 'prior_sample'), (9, 5, 32, 'normal'), (2, 16, 32, 'mean')])
def test_anytime_eval_vs_oracle(a3d_mod, weights, B, K, chunk, fill):
    ws = weights[('mn', 'trained')]
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=chunk)      # 9 x 5 decodes -> several chunks
    dec.set_weights(ws)
    rng = np.random.default_rng(B + K)
    z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
    mask = ar.bernoulli_mask(rng, B, 64, 0.5)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    tgt = ar.make_targets(rng, B)
    r = a3d_mod.anytime_eval(dec, z, mask, mu, tgt, K=K, seed=77, fill=fill, return_grid=True)
    zc = r['z_completed'].cpu().numpy()
    ref_mp, ref_cnt = ar.anytime_eval(dr.MODELNET_DECODER, ws, zc, tgt)      # oracle decodes the SAME completed latents
    mp = r['mean_prob'].cpu().numpy()
    assert np.abs(mp - ref_mp).max() < PROB_TOL
    nflip = int(((mp >= 0.5) != (ref_mp >= 0.5)).sum())
    assert nflip / mp.size < FLIP_TOL
    cnt = r['counts'].cpu().numpy()
    assert np.abs(cnt - ref_cnt).sum() <= 2 * nflip
    # exact invariants of the counts: TP + FN = target occupancy, TP + FP = predicted occupancy
    assert np.array_equal(cnt[:, 0] + cnt[:, 2], tgt.reshape(B, -1).sum(1).astype(np.int64))
    assert np.array_equal(cnt[:, 0] + cnt[:, 1], (mp.reshape(B, -1) >= 0.5).sum(1))
    # fused weighted-BCE loss: exact arithmetic check on the GPU's own mean grid, and close to the oracle's loss
    rl = a3d_mod.anytime_eval(dec, z, mask, mu, tgt, K=K, seed=77, fill=fill, return_loss=True, gamma=0.6)
    assert torch.equal(rl['counts'], r['counts'])
    np.testing.assert_allclose(rl['loss'].cpu().numpy(), ar.binary_loss(mp, tgt, gamma=0.6), rtol=1e-5)
    np.testing.assert_allclose(rl['loss'].cpu().numpy(), ar.binary_loss(ref_mp, tgt, gamma=0.6), rtol=2e-2)
    # counts-only call (no grid) and host-buffer call give the same integers
    r2 = a3d_mod.anytime_eval(dec, z, mask, mu, ar.pack_bits(tgt), K=K, seed=77, fill=fill)
    assert torch.equal(r2['counts'], r['counts'])
    from a3d import anytime_eval_host
    c3 = anytime_eval_host(dec, z, mask, mu, ar.pack_bits(tgt), K=K, seed=77, fill=fill)
    assert np.array_equal(c3, cnt)


def test_k_copies_of_one_latent_equal_single_decode(a3d_mod, decoders):
    dec = decoders[('mn', 'trained')]
    rng = np.random.default_rng(4)
    z1 = rng.standard_normal((2, 1, 64)).astype(np.float32)
    tgt = ar.make_targets(rng, 2)
    a = a3d_mod.anytime_eval(dec, None, None, None, tgt, z_completed=np.repeat(z1, 16, axis=1), return_grid=True)
    b = a3d_mod.anytime_eval(dec, None, None, None, tgt, z_completed=z1, return_grid=True)
    # the K-mean of 16 identical grids equals the single grid up to fp32 summation rounding
    assert (a['mean_prob'] - b['mean_prob']).abs().max().item() < 1e-6
    assert (a['counts'] - b['counts']).abs().sum().item() <= 4
    single = dec(z1[:, 0])
    assert np.array_equal(single, b['mean_prob'].cpu().numpy())


@pytest.mark.parametrize('B,K,chunk', [(37, 1, 32), (33, 1, 32), (3, 2, 32), (1, 1, 32), (2, 3, 32)])
def test_tail_kernel_matches_the_cuda_core_tail(a3d_mod, weights, B, K, chunk, monkeypatch):
    """The tcgen05 tail (tail_hcol.cu: 4 x 4 x 32 blocks, col2im in TMEM lanes / shuffles / shared memory) against the
    CUDA-core tail (tail.cu: direct 8-tap gather per output voxel) on the SAME 32^3 x 64 activations: A3D_TAIL_IMPL=simt
    keeps the tcgen05 hidden layers and swaps only the tail; the variable is read when the handle is created.  Odd batches
    and a ragged last chunk (32 + 5, 32 + 1 objects), K = 1 / even / odd."""
    ws = weights[('mn', 'trained')]
    rng = np.random.default_rng(100 + B)
    zc = rng.standard_normal((B, K, 64)).astype(np.float32)
    tgt = ar.make_targets(rng, B)
    monkeypatch.delenv('A3D_TAIL_IMPL', raising=False)
    new = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=chunk)
    monkeypatch.setenv('A3D_TAIL_IMPL', 'simt')
    old = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=chunk)
    monkeypatch.delenv('A3D_TAIL_IMPL')
    new.set_weights(ws)
    old.set_weights(ws)
    a = a3d_mod.anytime_eval(new, None, None, None, tgt, z_completed=zc, return_grid=True)
    b = a3d_mod.anytime_eval(old, None, None, None, tgt, z_completed=zc, return_grid=True)
    # same fp16 activations, fp16 x fp16 products are exact in fp32: only the fp32 summation order and the sigmoid form differ
    assert (a['mean_prob'] - b['mean_prob']).abs().max().item() < 2e-5
    nflip = int(((a['mean_prob'] >= 0.5) != (b['mean_prob'] >= 0.5)).sum().item())
    assert nflip <= 2
    assert (a['counts'] - b['counts']).abs().sum().item() <= 2 * nflip
    # counts-only path (tanh form of the sigmoid): same integers up to voxels whose mean sits within 1e-6 of the threshold
    c = a3d_mod.anytime_eval(new, None, None, None, ar.pack_bits(tgt), z_completed=zc)
    near = int(((a['mean_prob'] - 0.5).abs() < 1e-6).sum().item())
    assert (c['counts'] - a['counts']).abs().sum().item() <= 2 * near
    mp = a['mean_prob'].reshape(B, -1)
    assert torch.equal(a['counts'][:, 0] + a['counts'][:, 1], (mp >= 0.5).sum(1))


@pytest.mark.parametrize('n', [21, 40])
def test_l4_sweep_kernel_matches_the_generic_kernel(a3d_mod, weights, n, monkeypatch):
    """The 128->64 layer's w-sweep 2-CTA kernel (convt_l4_sw.cu: taps resolved in a TMEM accumulator ring) against the
    generic 1-CTA row-unit kernel of convt_tc.cu (A3D_L4_IMPL=generic, read at handle creation) on the same inputs: the
    32^3 x 64 activations agree to fp16 rounding of an fp32 sum taken in a different order (ragged n: 3 and 5 decode
    blocks of 8, i.e. a half-empty block pair and a partly filled block)."""
    ws = weights[('mn', 'trained')]
    z = torch.from_numpy(np.random.default_rng(n).standard_normal((n, 64)).astype(np.float32)).cuda()
    acts = {}
    for impl in (None, 'generic'):
        if impl:
            monkeypatch.setenv('A3D_L4_IMPL', impl)
        else:
            monkeypatch.delenv('A3D_L4_IMPL', raising=False)
        d = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=64)
        d.set_weights(ws)
        d(z)
        torch.cuda.synchronize()
        acts[impl] = d.debug_layer(4, n)
        d.close()
    monkeypatch.delenv('A3D_L4_IMPL', raising=False)
    a, b = acts[None], acts['generic']
    scale = max(1.0, float(np.abs(b).max()))
    assert np.abs(a - b).max() <= 2e-3 * scale       # one fp16 ulp of the largest activation
    assert (a == b).mean() > 0.98


def test_decode_and_eval_are_cuda_graph_capturable(a3d_mod, decoders):
    """every hot call is asynchronous on the caller's stream with no hidden synchronisation or
    allocation, so a decode (and a fused anytime evaluation) can be captured into a CUDA graph and replayed on new
    latents; the replay equals the eager call bit for bit."""
    dec = decoders[('mn', 'trained')]
    g0 = torch.Generator(device='cuda').manual_seed(5)
    z1 = torch.randn(8, 64, device='cuda', generator=g0)
    z2 = torch.randn(8, 64, device='cuda', generator=g0)
    tgt = ar.pack_bits(ar.make_targets(np.random.default_rng(3), 2))
    bits = torch.from_numpy(tgt).cuda()
    zin = z1.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # warm-up off the default stream, as torch asks before a capture
        dec(zin)
        a3d_mod.anytime_eval(dec, None, None, None, bits, z_completed=zin.view(2, 4, 64))
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = dec(zin)
        cnt = a3d_mod.anytime_eval(dec, None, None, None, bits, z_completed=zin.view(2, 4, 64))['counts']
    for z in (z2, z1):
        zin.copy_(z)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, dec(z))
        assert torch.equal(cnt, a3d_mod.anytime_eval(dec, None, None, None, bits, z_completed=z.view(2, 4, 64))['counts'])


def test_getEval_reference_return_tuple(a3d_mod, decoders):
    dec = decoders[('mn', 'trained')]
    rng = np.random.default_rng(8)
    B = 4
    z = rng.standard_normal((B, 64)).astype(np.float32)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    tgt = ar.make_targets(rng, B)
    cat = np.eye(40, dtype=np.float32)[rng.integers(0, 40, B)]
    out = a3d_mod.getEval(dec, (z, tgt, cat), mu, missing_prob=0.0)
    assert len(out) == 10 and out[5:] == (0, 0, 0, 0, 0) and tuple(out[0].shape) == (B, 64, 64, 64, 1)
    assert out[1] == pytest.approx(float(ar.binary_loss(out[0].cpu().numpy(), tgt, gamma=0.6).mean()), rel=1e-5)
    out = a3d_mod.getEval(dec, (z, tgt, cat), mu, missing_prob=0.5, K=4, seed=1, rng=np.random.default_rng(0))
    assert len(out) == 10 and tuple(out[5].shape) == (B, 64, 64, 64, 1) and 0.0 <= out[7] <= 1.0 and out[6] > 0


def test_full_size_properties_config2(a3d_mod, weights):
    """BASELINE config 2 size (256 objects x K=16 = 4096 decodes): size-independent exact invariants."""
    ws = weights[('mn', 'default')]
    dec = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=2048)       # two chunks of 128 objects
    dec.set_weights(ws)
    rng = np.random.default_rng(21)
    B, K = 256, 16
    z = rng.standard_normal((B, 64)).astype(np.float32)
    mask = ar.bernoulli_mask(rng, B, 64, 0.5)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    tgt8 = ar.make_targets(rng, 8)
    bits = np.tile(ar.pack_bits(tgt8), (B // 8, 1))
    r = a3d_mod.anytime_eval(dec, z, mask, mu, bits, K=K, seed=5)
    cnt = r['counts'].cpu().numpy()
    occ = np.tile(tgt8.reshape(8, -1).sum(1).astype(np.int64), B // 8)
    assert np.array_equal(cnt[:, 0] + cnt[:, 2], occ)                  # TP + FN = target occupancy, exactly
    assert (cnt >= 0).all() and (cnt.sum(1) <= 262144).all()
    # idempotence / determinism and shard invariance (objects 100..163 alone, with their global ids)
    r2 = a3d_mod.anytime_eval(dec, z, mask, mu, bits, K=K, seed=5)
    assert torch.equal(r2['counts'], r['counts'])
    r3 = a3d_mod.anytime_eval(dec, z[100:164], mask[100:164], mu, bits[100:164], K=K, seed=5, obj_offset=100)
    assert torch.equal(r3['counts'], r['counts'][100:164])
    # spot-check 2 objects of the big batch against the oracle
    zc = r['z_completed'][[3, 200]].cpu().numpy()
    tg = np.unpackbits(bits[[3, 200]], axis=1, bitorder='little').reshape(2, 64, 64, 64, 1)
    _, ref_cnt = ar.anytime_eval(dr.MODELNET_DECODER, ws, zc, tg)
    assert np.abs(cnt[[3, 200]] - ref_cnt).sum() <= 2 * FLIP_TOL * 262144 * 2


def test_prefix_sweep_properties_config4(a3d_mod, decoders):
    """BASELINE config 4 semantics on a small batch: prefix masks; the full prefix reproduces the plain decode and the
    sharded run (object offset) reproduces the slice of the unsharded one bit-for-bit."""
    dec = decoders[('mn', 'trained')]
    rng = np.random.default_rng(31)
    B = 6
    z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    tgt = ar.make_targets(rng, B)
    lengths = [1, 16, 32, 64]
    zz = np.repeat(z, len(lengths), axis=0)
    mask = np.tile(ar.prefix_mask(len(lengths), 64, lengths), (B, 1))
    bits = np.repeat(ar.pack_bits(tgt), len(lengths), axis=0)
    r = a3d_mod.anytime_eval(dec, zz, mask, mu, bits, K=1, seed=4, return_grid=True)
    zc = r['z_completed'].cpu().numpy()
    assert np.array_equal(zc[3::4, 0], z)                       # prefix 64 = nothing imputed
    full = dec(z)
    assert np.array_equal(r['mean_prob'].cpu().numpy()[3::4], full)
    assert np.array_equal(zc[0::4, 0, :1], z[:, :1]) and not np.array_equal(zc[0::4, 0, 1:], z[:, 1:])
    part = a3d_mod.anytime_eval(dec, zz[8:], mask[8:], mu, bits[8:], K=1, seed=4, obj_offset=8)
    assert torch.equal(part['counts'], r['counts'][8:])


def test_pascal_path_config3(a3d_mod, decoders, weights):
    """BASELINE config 3 decoder path: D = 16 latents drawn with sampling(mean, logvar), decoded and scored."""
    dec, ws = decoders[('pa', 'trained')], weights[('pa', 'trained')]
    rng = np.random.default_rng(32)
    B = 7
    mean = rng.standard_normal((B, 16)).astype(np.float32)
    logvar = np.clip(rng.standard_normal((B, 16)).astype(np.float32) - 2.0, -10, 10)
    z = a3d_mod.sampling(mean, logvar, seed=9, decoder=dec)
    eps = ar.philox_normals(9, np.arange(B, dtype=np.uint64), 1, 16)[:, 0]
    assert np.abs(z - ar.sampling(mean, logvar, eps)).max() < 2e-5
    tgt = ar.make_targets(rng, B)
    r = a3d_mod.anytime_eval(dec, None, None, None, tgt, z_completed=z[:, None, :], return_grid=True)
    ref_mp, ref_cnt = ar.anytime_eval(dr.PASCAL_DECODER, ws, z[:, None, :], tgt)
    mp = r['mean_prob'].cpu().numpy()
    assert np.isfinite(mp).all()
    nflip = int(((mp >= 0.5) != (ref_mp >= 0.5)).sum())
    assert np.abs(mp - ref_mp).max() < PROB_TOL and nflip / mp.size < FLIP_TOL
    assert np.abs(r['counts'].cpu().numpy() - ref_cnt).sum() <= 2 * nflip


@pytest.mark.parametrize('n', [1, 8, 9, 31, 33])
def test_ragged_decode_sizes_match_simt_path(a3d_mod, weights, n):
    """Every tile-size boundary of the tcgen05 kernels (8 / 16 / 32 / 128 decodes per tile, CTA pairs) against the
    CUDA-core diagnostic path on the same device."""
    ws = weights[('mn', 'trained')]
    z = np.random.default_rng(n).standard_normal((n, 64)).astype(np.float32)
    t = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=64)
    s = a3d_mod.decoder3D(dr.MODELNET_DECODER, max_chunk=64, impl='simt')
    t.set_weights(ws)
    s.set_weights(ws)
    a, b = t(z), s(z)
    assert np.isfinite(a).all() and np.abs(a - b).max() < 5e-3 and flips(a, b) < 2e-4


def test_getEval_values_match_the_oracle_composition(a3d_mod, decoders, weights):
    """nolbo.py:1472-1528 recomposed from oracle pieces (mean fill -> decode -> binary_loss(0.6) / precision / recall /
    nearest-prior accuracy; prior-sample fill with the same Philox draws -> K-mean decode -> the same metrics) against
    the 10-tuple a3d.getEval returns for an explicit mask."""
    dec, ws = decoders[('mn', 'trained')], weights[('mn', 'trained')]
    rng = np.random.default_rng(31)
    B, D, K = 4, 64, 2
    z = dr.round_bf16(rng.standard_normal((B, D)).astype(np.float32))
    mu = rng.standard_normal((40, D)).astype(np.float32)
    mask = ar.bernoulli_mask(rng, B, D, 0.5)
    tgt = ar.make_targets(rng, B)
    cat = np.eye(40, dtype=np.float32)[rng.integers(0, 40, B)]
    out = a3d_mod.getEval(dec, (z, tgt, cat), mu, missing_prob=0.5, K=K, seed=9, mask=mask)

    def metrics(prob, cnt, zc):
        c = cnt.astype(np.float64)
        pr = float((c[:, 0] / (c[:, 0] + c[:, 1] + 1e-10)).mean())                      # nolbo.py:1499-1501
        rc = float((c[:, 0] / (c[:, 0] + c[:, 2] + 1e-10)).mean())
        loss = float(ar.binary_loss(prob, tgt, gamma=0.6).mean())                        # :1497-1498
        acc = float((((zc[:, None, :] - mu[None]) ** 2).sum(-1).argmin(-1) == cat.argmax(-1)).mean())   # :1488-1494
        return loss, pr, rc, acc

    z_mean, _ = ar.impute(z, mask, mu, 1, seed=9, fill='mean')                           # :1477-1482
    p0, c0 = ar.anytime_eval(dr.MODELNET_DECODER, ws, z_mean, tgt)
    zc, _ = ar.impute(z, mask, mu, K, seed=9, fill='prior_sample')                       # :1505-1510 (K draws)
    p1, c1 = ar.anytime_eval(dr.MODELNET_DECODER, ws, zc, tgt)
    for got_prob, got, (prob, cnt, zz) in ((out[0], out[1:5], (p0, c0, z_mean[:, 0])), (out[5], out[6:10], (p1, c1, zc[:, 0]))):
        loss, pr, rc, acc = metrics(prob, cnt, zz)
        assert np.abs(got_prob.cpu().numpy() - prob).max() < PROB_TOL
        assert got[0] == pytest.approx(loss, rel=2e-3)
        assert got[1] == pytest.approx(pr, abs=2e-3) and got[2] == pytest.approx(rc, abs=2e-3)
        assert got[3] == pytest.approx(acc, abs=1e-9)
