"""Small-batch and host-return measurements (diagnostic): the reference's own call shapes.
  * decoder(z) at 32 / 72 decodes per call (nolbo_test.py:167-177, test_modelnet_VAE_dr.py:52): eager vs CUDA-graph
    replay, device in / device out, plus per-stage device times;
  * anytime_eval at B=1 x K=32 and B=32 x K=1 (counts only);
  * decoder(z) numpy -> numpy through a3d_decode_host at B=72 / 4096, fp32 / fp16 / bit outputs, pinned output buffer,
    against the measured pinned D2H rate of the box.
Usage: python tests/tools/bench_small.py [--big 4096] [--small-only]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr

big = int(sys.argv[sys.argv.index('--big') + 1]) if '--big' in sys.argv else 4096
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=256)
dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))


def timed(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3      # us


for B in (32, 72):
    z = torch.randn(B, 64, device='cuda')
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dec(z)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = dec(z)
    te, tg = timed(lambda: dec(z)), timed(g.replay)
    dec.set_profiling(True)
    rec = []
    for _ in range(6):
        dec(z)
        rec.append(dec.stage_times_ms())
    dec.set_profiling(False)
    med = {k: round(float(np.median([r[k] for r in rec[1:]])) * 1e3, 1) for k in rec[0]}
    print(f'decoder(z) B={B}: eager {te:.1f} us, graph replay {tg:.1f} us = {B * 6.66383e9 / (tg * 1e-6) / 1e12:.0f} TFLOP/s; '
          f'stage us {med} sum {sum(med.values()):.1f}', flush=True)

for B, K in ((1, 32), (32, 1)):
    zc = torch.randn(B, K, 64, device='cuda')
    bits = torch.zeros((B, 32768), dtype=torch.uint8, device='cuda')
    t = timed(lambda: a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc))
    print(f'anytime_eval B={B} K={K} (counts only, device in/out): {t:.1f} us per call', flush=True)

if '--small-only' in sys.argv:
    sys.exit(0)

# pinned D2H rate of this box
src = torch.empty(1 << 28, dtype=torch.uint8, device='cuda')
dst = torch.empty(1 << 28, dtype=torch.uint8, pin_memory=True)
for _ in range(2):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
d2h = 4 * (1 << 28) / (time.perf_counter() - t0) / 1e9
print(f'pinned D2H copy rate: {d2h:.1f} GB/s', flush=True)
del src, dst

big_dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=256)
big_dec.set_weights(dec.get_weights())
for B in (72, big):
    z = np.random.default_rng(B).standard_normal((B, 64)).astype(np.float32)
    for dt, per in (('f32', 262144 * 4), ('f16', 262144 * 2), ('bits', 32768)):
        shape = (B, 32768) if dt == 'bits' else (B, 64, 64, 64, 1)
        out = a3d.pinned_empty(shape, {'f32': np.float32, 'f16': np.float16, 'bits': np.uint8}[dt])
        reps = 3 if B > 1000 else 20
        for _ in range(2):
            big_dec(z, out=out, out_dtype=dt)
        t0 = time.perf_counter()
        for _ in range(reps):
            big_dec(z, out=out, out_dtype=dt)
        dt_s = (time.perf_counter() - t0) / reps
        bound = B * per / (d2h * 1e9)
        print(f'decode_host B={B} {dt}: {dt_s * 1e3:.2f} ms per call = {B / dt_s:.0f} decodes/s; D2H bound {bound * 1e3:.2f} ms '
              f'-> {100 * bound / dt_s:.0f} % of the PCIe bound', flush=True)
    if B <= 100:
        t0 = time.perf_counter()
        for _ in range(5):
            big_dec(z)
        print(f'decode_host B={B} f32 into a fresh pageable array: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call', flush=True)
