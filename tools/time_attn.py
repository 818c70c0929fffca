"""Time the attention kernels alone (CUDA events) at the workload's shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import ops

def run(Bq, N, C, heads, impl, reps=20, vpad=0):
    R, d = 4, C // heads
    ld = 3 * C if not vpad else 2 * C + heads * vpad
    qkv = torch.randn((Bq, N, ld), device="cuda").half()
    refkv = torch.randn((2 * R, N, ld), device="cuda").half()
    if vpad:
        for t in (qkv, refkv):
            t[..., 2 * C:].reshape(t.shape[0], N, heads, vpad)[..., d] = 1.0
    F = Bq // 2
    rows = [[h * F + f] + [-(h * R + r) - 1 for r in range(4)] for h in range(2) for f in range(F)]
    idx = torch.tensor(rows, dtype=torch.int32, device="cuda")
    ops.set_attn_impl(impl)
    call = lambda: ops.attention(qkv, 0, ld, qkv, C, 2 * C, ld, refkv, C, 2 * C, ld, Bq, N, N, heads, d, idx, [0.6, .1, .1, .1, .1], v_head_stride=(vpad or d))
    try:
        for _ in range(3): call()
    except Exception as e:
        ops.set_attn_impl(0); return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ops.set_attn_impl(0)
    t = e0.elapsed_time(e1) / reps / 1e3
    return t, Bq * 5 * 4.0 * N * N * C / t / 1e12

for (Bq, N, C) in [(8, 4096, 320), (24, 4096, 320), (8, 1024, 640), (24, 1024, 640), (24, 256, 1280)]:
    for impl, name in ((1, "tcgen05"), (2, "mma.sync")):
        r = run(Bq, N, C, 8, impl)
        if r: print(f"B={Bq} N={N} C={C} {name}: {r[0]*1e3:.3f} ms  {r[1]:.1f} TFLOP/s", flush=True)
        if impl == 1 and C == 320:
            r = run(Bq, N, C, 8, impl, vpad=48)
            if r: print(f"B={Bq} N={N} C={C} {name} ones-column: {r[0]*1e3:.3f} ms  {r[1]:.1f} TFLOP/s", flush=True)
