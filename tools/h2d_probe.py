"""H2D diagnostics for the e2e path: peak pinned copy bandwidth vs the per-frame upload calls (run on the GPU box)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libviso_b200 import api, build, synth
build.build()
torch.cuda.set_device(0)
big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for n in (1 << 30, 1 << 24, 1 << 20, 466616, 1 << 16):
    reps = max(1, (1 << 30) // n)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(reps):
        dev[i * n:(i + 1) * n].copy_(big[i * n:(i + 1) * n], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"copy size {n:>10d} x {reps:5d}: {reps * n / dt / 1e9:6.1f} GB/s  ({dt / reps * 1e6:8.1f} us per copy)")
F = 1000
ctx = api.Context(0); ctx.set_image_extent(synth.W, synth.H)
seq = ctx.sequence(F, 2100, 121, 50); seq.set_image_size(synth.W, synth.H)
im = torch.zeros(synth.H * synth.W, dtype=torch.uint8).pin_memory()
kp = torch.zeros(2040 * 2, dtype=torch.float32).pin_memory()
for rep in range(3):
    ctx.sync(); t = time.perf_counter()
    for f in range(F):
        seq.upload_frame_images_raw(f, im.data_ptr(), im.data_ptr(), kp.data_ptr(), 2040, kp.data_ptr(), 2040)
    t_enq = time.perf_counter() - t
    ctx.sync(); dt = time.perf_counter() - t
    print(f"upload 1000 frames (images): enqueue {t_enq * 1e3:.1f} ms, done {dt * 1e3:.1f} ms, {F * (2 * synth.W * synth.H + 2 * 2040 * 8) / dt / 1e9:.1f} GB/s")
