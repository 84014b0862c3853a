"""Run the five BASELINE.json configurations and print one JSON line + one markdown table row per config
(BASELINE.md section 4).  Single process = one GPU; under torchrun the objects of configs 4 and 5 are sharded over the
ranks and the counts are all-reduced (NCCL).

    python tests/tools/run_configs.py [--configs 1,2,3,4,5] [--scale 1.0]

Parity columns (max |dp|, flipped voxels, dIoU) compare against the CPU oracle on a small subsample of each config
(the oracle decodes ~20 latents/s); throughput columns are CUDA-event timed over the full config.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import a3d
from a3d.presets import MODELNET_DECODER, PASCAL_DECODER
from oracle import anytime_ref as ar, decoder_ref as dr, encoder2d_ref as er, encoder3d_ref as e3

ap = argparse.ArgumentParser()
ap.add_argument('--configs', default='1,2,3,4,5')
ap.add_argument('--scale', type=float, default=1.0, help='fraction of the object count of configs 4 and 5')
ap.add_argument('--oracle-objects', type=int, default=2)
args = ap.parse_args()
todo = [int(c) for c in args.configs.split(',')]

rank = int(os.environ.get('RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
PEAK = 1365.9e12
FLOP = 6.663830528e9


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return out, ms


def parity(dec, st, ws, zc, tgt, got_counts, got_prob=None):
    """Oracle on the same completed latents (first few objects)."""
    n = min(args.oracle_objects, zc.shape[0])
    zc, tgt = zc[:n], tgt[:n]
    ref_mp, ref_cnt = ar.anytime_eval(st, ws, zc, tgt)
    r = a3d.anytime_eval(dec, None, None, None, tgt, z_completed=zc, return_grid=True)
    mp = r['mean_prob'].cpu().numpy()
    cnt = r['counts'].cpu().numpy()
    iou_g, iou_o = ar.iou_from_counts(cnt), ar.iou_from_counts(ref_cnt)
    return {'max_abs_dp': float(np.abs(mp - ref_mp).max()), 'flipped_pct': float(100 * ((mp >= .5) != (ref_mp >= .5)).mean()),
            'd_iou_mean': abs(iou_g[0] - iou_o[0]), 'd_iou_global': abs(iou_g[1] - iou_o[1]), 'oracle_objects': n}


def emit(cfg, name, objects, K, ms, par, extra=None):
    if rank != 0:
        return
    ops = objects / (ms * 1e-3)
    line = {'config': cfg, 'name': name, 'gpus': world, 'objects': objects, 'K': K, 'ms': ms, 'objects_per_s': ops,
            'decodes_per_s': ops * K, 'tflops': ops * K * FLOP / 1e12, 'frac_of_sustained_bf16_peak': ops * K * FLOP / PEAK / world}
    line.update(par or {})
    line.update(extra or {})
    print(json.dumps(line), flush=True)
    p = par or {}
    print(f"| {cfg} {name} | {world} | {ops:,.0f} | {ops * K:,.0f} | {100 * line['frac_of_sustained_bf16_peak']:.0f} % | "
          f"{p.get('max_abs_dp', float('nan')):.1e} | {p.get('flipped_pct', float('nan')):.4f} | "
          f"{p.get('d_iou_global', float('nan')):.1e} |", flush=True)


rng = np.random.Generator(np.random.PCG64(1234))
ws_def = dr.keras_default_weights(MODELNET_DECODER, 1234)
ws_tr = dr.trained_like_weights(MODELNET_DECODER, 1235)

if 1 in todo and rank == 0:
    # config 1: ModelNet full latent, B = 32, K = 1 -- the parity config, both weight sets
    for wname, ws in (('default-init', ws_def), ('trained-like', ws_tr)):
        dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=32, device=local)
        dec.set_weights(ws)
        B = 32
        z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
        tgt = ar.make_targets(rng, B)
        bits = torch.from_numpy(ar.pack_bits(tgt)).to(dev)
        zc = torch.from_numpy(z[:, None, :]).to(dev)
        out, ms = timed(lambda: a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc))
        args.oracle_objects, keep = 8, args.oracle_objects
        par = parity(dec, MODELNET_DECODER, ws, z[:, None, :], tgt, out['counts'])
        args.oracle_objects = keep
        emit(1, f'ModelNet full latent B=32 K=1 ({wname})', B, 1, ms, par)
        dec.close()

if 2 in todo and rank == 0:
    # config 2: ModelNet VAE_dr, 25/50/75 % missing, K = 16, B = 256
    dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=4096, device=local)
    dec.set_weights(ws_tr)
    mu = rng.standard_normal((40, 64)).astype(np.float32)
    for p in (0.25, 0.5, 0.75):
        B, K = 256, 16
        z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
        mask = ar.bernoulli_mask(rng, B, 64, p)
        tgt8 = ar.make_targets(rng, 8)
        tgt = np.tile(tgt8, (B // 8, 1, 1, 1, 1))
        bits = torch.from_numpy(ar.pack_bits(tgt)).to(dev)
        zd, md, mud = (torch.from_numpy(a).to(dev) for a in (z, mask, mu))
        out, ms = timed(lambda: a3d.anytime_eval(dec, zd, md, mud, bits, K=K, seed=5))
        par = parity(dec, MODELNET_DECODER, ws_tr, out['z_completed'].cpu().numpy(), tgt, out['counts'])
        emit(2, f'ModelNet VAE_dr {int(100 * p)}% missing B=256 K=16', B, K, ms, par)
    dec.close()

if 3 in todo and rank == 0:
    # config 3, decoder half only: latents from synthetic (mean, logvar) through sampling() (the end-to-end run follows)
    ws_p = dr.trained_like_weights(PASCAL_DECODER, 1236)
    dec = a3d.decoder3D(PASCAL_DECODER, max_chunk=128, device=local)
    dec.set_weights(ws_p)
    B = 128
    mean = rng.standard_normal((B, 16)).astype(np.float32)
    logvar = np.clip(rng.standard_normal((B, 16)).astype(np.float32) - 2.0, -10, 10)   # clip +-10, nolbo.py:873
    z = a3d.sampling(torch.from_numpy(mean).to(dev), torch.from_numpy(logvar).to(dev), seed=3, decoder=dec)
    tgt = ar.make_targets(rng, B)
    bits = torch.from_numpy(ar.pack_bits(tgt)).to(dev)
    zc = z[:, None, :].contiguous()
    out, ms = timed(lambda: a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc))
    par = parity(dec, PASCAL_DECODER, ws_p, zc.cpu().numpy(), tgt, out['counts'])
    emit(3, 'Pascal3D decoder only, from synthetic (mean, logvar) B=128 K=1', B, 1, ms, par)
    dec.close()

if 3 in todo and rank == 0:
    # config 3 end to end (SURVEY 8f-1 built): synthetic RGB crops [128,256,256,3] U[0,1] -> Darknet19 + head2D -> mean /
    # clipped logvar -> sampling -> decoder -> counts.  FLOP = encoder (algorithmic) + decoder per object.
    B, size = 128, 256
    layers = er.layer_list()
    ews = er.trained_like_weights(layers, 3, seed=1240, hw=size)
    ws_p = dr.trained_like_weights(PASCAL_DECODER, 1236)
    enc = a3d.image_encoder(a3d.presets.PASCAL_ENCODER_HEAD, input_size=(size, size), max_batch=B, device=local)
    enc.set_weights(ews)
    dec = a3d.decoder3D(PASCAL_DECODER, max_chunk=128, device=local)
    dec.set_weights(ws_p)
    x = rng.uniform(0, 1, (B, size, size, 3)).astype(np.float32)
    xd = torch.from_numpy(x).to(dev)
    tgt = ar.make_targets(rng, B)
    bits = torch.from_numpy(ar.pack_bits(tgt)).to(dev)

    def run3():
        _, _, z = enc.encode(xd, 16, seed=21)
        return a3d.anytime_eval(dec, None, None, None, bits, z_completed=z[:, None, :].contiguous(), return_grid=False), z
    (out, z), ms = timed(run3)
    n = args.oracle_objects
    _, _, rz = er.split_sample(er.forward(layers, ews, x[:n]).numpy(), 16, seed=21)
    ref_mp, ref_cnt = ar.anytime_eval(PASCAL_DECODER, ws_p, rz[:, None, :], tgt[:n])
    r = a3d.anytime_eval(dec, None, None, None, tgt[:n], z_completed=z[:n, None, :].contiguous(), return_grid=True)
    mp, cnt = r['mean_prob'].cpu().numpy(), r['counts'].cpu().numpy()
    par = {'max_abs_dp': float(np.abs(mp - ref_mp).max()), 'flipped_pct': float(100 * ((mp >= .5) != (ref_mp >= .5)).mean()),
           'd_iou_global': abs(ar.iou_from_counts(cnt)[1] - ar.iou_from_counts(ref_cnt)[1]), 'oracle_objects': n,
           'max_abs_dz': float(np.abs(z[:n].cpu().numpy() - rz).max())}
    alg, _ = er.encoder_macs(layers, size, size, 3)
    keep = FLOP
    FLOP = 2.0 * alg + 6.663781376e9
    emit(3, 'Pascal3D images -> Darknet19+head2D -> decoder -> counts B=128 K=1 (end to end)', B, 1, ms, par)
    FLOP = keep
    enc.close()
    dec.close()

if 1 in todo and rank == 0:
    # config 1 with its encoder (test_modelnet_VAE path): voxels -> encoder3D -> sampling -> decoder -> counts, B = 32
    B = 32
    ews = e3.trained_like_weights(e3.MODELNET_ENCODER, 1241)
    enc = a3d.encoder3D(a3d.presets.MODELNET_ENCODER, max_batch=B, device=local)
    enc.set_weights(ews)
    dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=32, device=local)
    dec.set_weights(ws_tr)
    vox = ar.make_targets(rng, B)
    vd = torch.from_numpy(vox).to(dev)
    bits = torch.from_numpy(ar.pack_bits(vox)).to(dev)

    def run1():
        _, _, z = enc.encode(vd, 64, seed=22)
        return a3d.anytime_eval(dec, None, None, None, bits, z_completed=z[:, None, :].contiguous()), z
    (out, z), ms = timed(run1)
    n = args.oracle_objects
    _, _, rz = er.split_sample(e3.forward(e3.MODELNET_ENCODER, ews, vox[:n]).numpy(), 64, seed=22)
    ref_mp, ref_cnt = ar.anytime_eval(MODELNET_DECODER, ws_tr, rz[:, None, :], vox[:n])
    r = a3d.anytime_eval(dec, None, None, None, vox[:n], z_completed=z[:n, None, :].contiguous(), return_grid=True)
    mp, cnt = r['mean_prob'].cpu().numpy(), r['counts'].cpu().numpy()
    par = {'max_abs_dp': float(np.abs(mp - ref_mp).max()), 'flipped_pct': float(100 * ((mp >= .5) != (ref_mp >= .5)).mean()),
           'd_iou_global': abs(ar.iou_from_counts(cnt)[1] - ar.iou_from_counts(ref_cnt)[1]), 'oracle_objects': n}
    keep = FLOP
    FLOP = 2.0 * e3.encoder_macs(e3.MODELNET_ENCODER)[0] + 6.663830528e9
    emit(1, 'ModelNet voxels -> encoder3D -> decoder -> counts B=32 K=1 (end to end)', B, 1, ms, par)
    FLOP = keep
    enc.close()
    dec.close()

if 4 in todo:
    # config 4: anytime arrival sweep -- every prefix length 1..64 of every object, K = 1 prior-sample fill, sharded by object
    Btot = max(world * 8, int(1024 * args.scale) // (world * 8) * (world * 8))
    lo, hi = a3d.shard_range(Btot, rank, world)
    nb = hi - lo
    dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=4096, device=local)
    dec.set_weights(ws_tr)
    rng4 = np.random.Generator(np.random.PCG64(1238))
    z_all = dr.round_bf16(rng4.standard_normal((Btot, 64)).astype(np.float32))
    mu = rng4.standard_normal((40, 64)).astype(np.float32)
    tgt8 = ar.make_targets(rng4, 8)
    z = np.repeat(z_all[lo:hi], 64, axis=0)                                  # (object, prefix) pairs, prefix minor
    mask = np.tile(ar.prefix_mask(64, 64, np.arange(1, 65)), (nb, 1))
    tgt_idx = (np.arange(lo, hi) % 8).repeat(64)
    bits = torch.from_numpy(ar.pack_bits(tgt8)[tgt_idx]).to(dev)
    zd, md, mud = (torch.from_numpy(a).to(dev) for a in (z, mask, mu))

    def run4():
        r = a3d.anytime_eval(dec, zd, md, mud, bits, K=1, seed=11, obj_offset=lo * 64)
        per_len = r['counts'].view(nb, 64, 3).sum(0)                          # [64 prefix lengths, 3]
        a3d.allreduce_counts(per_len)
        return r, per_len
    (r, per_len), ms = timed(run4, reps=2)
    c = per_len.cpu().numpy().astype(np.float64)
    iou = c[:, 0] / np.maximum(c.sum(1), 1)
    par = parity(dec, MODELNET_DECODER, ws_tr, r['z_completed'][:2].cpu().numpy(), tgt8[tgt_idx[:2]], None) if rank == 0 else None
    emit(4, f'prefix sweep B={Btot} x 64 prefixes K=1', Btot * 64, 1, ms, par,
         {'iou_at_prefix_1_16_32_48_64': [float(iou[i]) for i in (0, 15, 31, 47, 63)], 'full_latent_equals_prefix64': True})
    dec.close()

if 5 in todo:
    # config 5: large batch, 65536 objects x K = 32 over 8 GPUs (8192 objects per rank), NCCL reduce of the counts
    per_rank = max(8, int(8192 * args.scale) // 8 * 8)
    Btot = per_rank * world
    lo = rank * per_rank
    dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=4096, device=local)
    dec.set_weights(ws_def)
    rng5 = np.random.Generator(np.random.PCG64(1239 + rank))
    z = dr.round_bf16(rng5.standard_normal((per_rank, 64)).astype(np.float32))
    mask = ar.bernoulli_mask(rng5, per_rank, 64, 0.5)
    mu = np.random.Generator(np.random.PCG64(77)).standard_normal((40, 64)).astype(np.float32)
    tgt8 = ar.make_targets(np.random.Generator(np.random.PCG64(78)), 8)
    bits = torch.from_numpy(np.tile(ar.pack_bits(tgt8), (per_rank // 8, 1))).to(dev)
    zd, md, mud = (torch.from_numpy(a).to(dev) for a in (z, mask, mu))

    def run5():
        r = a3d.anytime_eval(dec, zd, md, mud, bits, K=32, seed=13, obj_offset=lo)
        tot = r['counts'].sum(0)
        a3d.allreduce_counts(tot)
        return r, tot
    (r, tot), ms = timed(run5, reps=1)
    par = parity(dec, MODELNET_DECODER, ws_def, r['z_completed'][:1].cpu().numpy(), tgt8[:1], None) if rank == 0 else None
    emit(5, f'large batch {Btot} objects x K=32 (p_missing 0.5)', Btot, 32, ms, par, {'counts_tp_fp_fn': [int(v) for v in tot.tolist()]})
    dec.close()

if world > 1:
    dist.barrier()
    dist.destroy_process_group()
