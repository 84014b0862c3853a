"""The fused tail timed ALONE (a3d_debug_time_tail: the tail kernel back to back on the activations of one 4096-decode
chunk) next to its time inside a full step, with the SM clock sampled during each phase: does the tail's HBM rate follow
the clock the power cap leaves it?  Usage: python tests/tools/tail_alone.py [reps]"""
import ctypes as C, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B, K = 256, 16
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=B * K)
dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))
zc = torch.randn(B, K, 64, device='cuda')
bits = torch.zeros((B, 32768), dtype=torch.uint8, device='cuda')
uuid = 'GPU-' + str(torch.cuda.get_device_properties(0).uuid)


class Clocks:
    def __enter__(self):
        self.lines = []
        self.p = subprocess.Popen(['nvidia-smi', '-i', uuid, '--query-gpu=clocks.sm,clocks.mem,power.draw', '--format=csv,noheader,nounits',
                                   '-lms', '50'], stdout=subprocess.PIPE, text=True)
        self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True)
        self.t.start()
        time.sleep(1.0)
        self.n0 = len(self.lines)
        return self

    def __exit__(self, *a):
        time.sleep(0.1)
        self.p.terminate()
        v = np.array([[float(x) for x in l.split(',')] for l in self.lines[self.n0:] if l.count(',') == 2])
        self.sm, self.mem, self.pw = (np.median(v[:, i]) for i in range(3)) if len(v) else (0, 0, 0)


tail_bytes = B * (K * 2_097_152 * 2 + 32768 + 24)
# full steps (the tail inside the chain, library profiling mode)
for _ in range(3):
    a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc)
dec.set_profiling(True)
with Clocks() as c:
    t = []
    t0 = time.time()
    while time.time() - t0 < 3.0:
        a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc)
        t.append(dec.stage_times_ms()['tail'])
dec.set_profiling(False)
ms = float(np.median(t))
print(f'tail inside full steps : {ms:.3f} ms = {tail_bytes / ms / 1e6:.0f} GB/s at SM {c.sm:.0f} MHz, mem {c.mem:.0f} MHz, {c.pw:.0f} W', flush=True)
# the tail alone, back to back
counts = torch.zeros((B, 3), dtype=torch.int64, device='cuda')
out = C.c_float(0)
with Clocks() as c:
    for _ in range(6):
        a3d._capi.check(dec._lib.a3d_debug_time_tail(dec._h, B, K, bits.data_ptr(), counts.data_ptr(), reps, C.byref(out), 0), 'time_tail')
ms = out.value
print(f'tail alone, back to back: {ms:.3f} ms = {tail_bytes / ms / 1e6:.0f} GB/s at SM {c.sm:.0f} MHz, mem {c.mem:.0f} MHz, {c.pw:.0f} W', flush=True)
# a plain device-to-device copy of the same bytes for reference (read + write)
src = torch.empty(1 << 31, dtype=torch.uint8, device='cuda')
dst = torch.empty_like(src)
with Clocks() as c:
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        dst.copy_(src)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print(f'torch copy 2 GiB        : {ms:.3f} ms = {2 * (1 << 31) / ms / 1e6:.0f} GB/s (read + write) at SM {c.sm:.0f} MHz, {c.pw:.0f} W', flush=True)
