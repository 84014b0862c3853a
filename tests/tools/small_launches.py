"""decoder(z) at 32 and 72 latents, three calls each: the command behind the ncu launch list of the small-call shapes
(per-kernel device times without launch gaps).  Usage: ncu --metrics gpu__time_duration.sum ... python tests/tools/small_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=96)
dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))
for B in (32, 72):
    z = torch.randn(B, 64, device='cuda')
    for _ in range(3):
        dec(z)
    torch.cuda.synchronize()
print('ok')
