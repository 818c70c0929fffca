"""The dominant kernel alone at the workload's shape and LAYOUT (fused q|k|v projection whose V heads are padded to 48
columns with a ones column, as SD15Denoiser lays it out): N=4096, d=40, self + 4 cached refs, GCB_PROFILE_BQ CFG rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import ops
Bq, N, C, R, VS = int(os.environ.get("GCB_PROFILE_BQ", "24")), 4096, 320, 4, int(os.environ.get("GCB_PROFILE_VSTRIDE", "48"))
ld = 2 * C + 8 * VS
qkv = torch.randn((Bq, N, ld), device="cuda").half()
refkv = torch.randn((2 * R, N, ld), device="cuda").half()
if VS > 40:
    for t in (qkv, refkv):
        v = t[..., 2 * C:].reshape(t.shape[0], N, 8, VS)
        v[..., 40:] = 0
        v[..., 40] = 1.0
rows = [[h * (Bq // 2) + f] + [-(h * R + r) - 1 for r in range(4)] for h in range(2) for f in range(Bq // 2)]
idx = torch.tensor(rows, dtype=torch.int32, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = ops.attention(qkv, 0, ld, qkv, C, 2 * C, ld, refkv, C, 2 * C, ld, Bq, N, N, 8, 40, idx, [0.6, .1, .1, .1, .1],
                        v_head_stride=VS)
torch.cuda.synchronize()
print(float(out.float().abs().mean()))
