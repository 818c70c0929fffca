"""GroupNorm / LayerNorm alone at the workload's shapes (CUDA events): time and HBM bytes (GN: 2 reads + 1 write of the
tensor, LN: 1 read + 1 write) against the copy bandwidth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import ops


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (B, H, C, silu) in [(72, 64, 320, True), (72, 64, 640, True), (72, 32, 640, True), (72, 32, 1280, True),
                        (72, 16, 1280, True), (72, 8, 1280, True), (4, 512, 128, True), (4, 256, 256, True)]:
    x = torch.randn((B, H, H, C), device="cuda").half()
    g, b = torch.randn(C, device="cuda").half(), torch.randn(C, device="cuda").half()
    us = timeit(lambda: ops.groupnorm(x, None, g, b, 32, 1e-5, silu))
    byts = x.numel() * 2 * 3
    print(f"groupnorm B={B} {H}x{H} C={C}: {us:8.1f} us  {byts / us / 1e6:6.2f} TB/s (3 passes)", flush=True)
for (M, C) in [(72 * 4096, 320), (72 * 1024, 640), (72 * 256, 1280)]:
    x = torch.randn((M, C), device="cuda").half()
    g, b = torch.randn(C, device="cuda").half(), torch.randn(C, device="cuda").half()
    us = timeit(lambda: ops.layernorm(x, g, b))
    print(f"layernorm M={M} C={C}: {us:8.1f} us  {x.numel() * 4 / us / 1e6:6.2f} TB/s (2 passes)", flush=True)
