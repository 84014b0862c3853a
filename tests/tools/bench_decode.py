import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr
B = 256
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=B)
dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))
z = torch.randn(B, 64, device='cuda')
for _ in range(2): out = dec(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = dec(z)
e1.record(); torch.cuda.synchronize()
print('decoder(z) device in/out: %.3f ms per %d decodes = %.0f decodes/s' % (e0.elapsed_time(e1)/5, B, B*5/(e0.elapsed_time(e1)*1e-3)), type(out), getattr(out,'shape',None))

# small batch (config 1: B = 32): eager launches vs one CUDA-graph replay of the same call
B2 = 32
z32 = torch.randn(B2, 64, device='cuda')
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    dec(z32)
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    o32 = dec(z32)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
te, tg = t(lambda: dec(z32)), t(g.replay)
print('decoder(z) B = 32: eager %.1f us (%.0f decodes/s), CUDA-graph replay %.1f us (%.0f decodes/s)' % (te * 1e6, B2 / te, tg * 1e6, B2 / tg))
