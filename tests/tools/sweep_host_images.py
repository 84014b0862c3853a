import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, a3d
from oracle import decoder_ref as dr, encoder2d_ref as er
B, size, D16 = 128, 256, 16
rng = np.random.default_rng(3)
x_host = rng.random((B, size, size, 3), dtype=np.float32)
x_pin = torch.from_numpy(x_host).pin_memory()
x_u8 = torch.from_numpy(np.rint(x_host * 255).astype(np.uint8)).pin_memory()
dec = a3d.decoder3D(a3d.presets.PASCAL_DECODER, max_chunk=B)
dec.set_weights(dr.keras_default_weights(a3d.presets.PASCAL_DECODER, 1))
bits = torch.zeros((B, 32768), dtype=torch.uint8, device='cuda')
ones = torch.ones((B, D16), device='cuda')
layers = er.layer_list()
ws = er.keras_default_weights(layers, 3, seed=1)
for mb in (32, 64, 128):
    enc = a3d.image_encoder(a3d.presets.PASCAL_ENCODER_HEAD, input_size=(size, size), max_batch=mb)
    enc.set_weights(ws)
    for name, x in (('f32', x_pin), ('u8', x_u8)):
        def step(i):
            _, _, z = enc.encode(x, D16, seed=100 + i)
            return a3d.anytime_eval(dec, z, ones, None, bits, K=1, seed=i, fill='normal')['counts']
        for i in range(3): step(i).cpu()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(10): step(i).cpu()
        dt = (time.perf_counter() - t0) / 10
        print(f'max_batch {mb} {name}: {dt*1e3:.3f} ms per {B} -> {B/dt:.0f} objects/s', flush=True)
    enc.close()
