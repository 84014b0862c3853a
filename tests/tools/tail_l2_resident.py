"""How fast (and at what board power) does the fused tail run when its 32^3 x 64 activations are L2 resident?  The tail
kernel alone, back to back (a3d_debug_time_tail), on 16 / 24 / 32 / 64 / 256 decodes (64 MB ... 1 GB of activations; the
L2 holds 126 MB), with the SM clock and power sampled.  Feasibility number for streaming the activations from the 128->64
kernel to the tail through L2.  Usage: python tests/tools/tail_l2_resident.py"""
import ctypes as C, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr

uuid = 'GPU-' + str(torch.cuda.get_device_properties(0).uuid)
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=256)
dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))


def sample_clocks(fn):
    lines = []
    p = subprocess.Popen(['nvidia-smi', '-i', uuid, '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '50'],
                         stdout=subprocess.PIPE, text=True)
    t = threading.Thread(target=lambda: [lines.append(l) for l in p.stdout], daemon=True)
    t.start()
    time.sleep(0.8)
    n0 = len(lines)
    r = fn()
    time.sleep(0.1)
    p.terminate()
    v = np.array([[float(x) for x in l.split(',')] for l in lines[n0:] if l.count(',') == 1])
    return r, (np.median(v[:, 0]), np.median(v[:, 1])) if len(v) else (0, 0)


for n in (16, 24, 32, 64, 256):
    K = 8
    B = n // K
    zc = torch.randn(B, K, 64, device='cuda')
    bits = torch.zeros((B, 32768), dtype=torch.uint8, device='cuda')
    a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc)
    counts = torch.zeros((B, 3), dtype=torch.int64, device='cuda')
    out = C.c_float(0)
    reps = max(200, 200000 // n)

    def run():
        for _ in range(3):
            a3d._capi.check(dec._lib.a3d_debug_time_tail(dec._h, B, K, bits.data_ptr(), counts.data_ptr(), reps, C.byref(out), 0), 'time_tail')
        return out.value
    ms, (sm, pw) = sample_clocks(run)
    gb = n * 4.194304e6 / (ms * 1e-3) / 1e9
    print(f'{n:4d} decodes ({n * 4.19:.0f} MB): {ms * 1e3 / n:.3f} us per decode = {gb:.0f} GB/s of activations, SM {sm:.0f} MHz, {pw:.0f} W', flush=True)
