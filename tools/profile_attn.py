"""The dominant kernel alone at the workload's shape (view batch: 6 CFG rows, N=4096, d=40, self + 4 cached refs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import ops
Bq, N, C, R = int(os.environ.get("GCB_PROFILE_BQ", "24")), 4096, 320, 4
qkv = torch.randn((Bq, N, 3 * C), device="cuda").half()
refkv = torch.randn((2 * R, N, 3 * C), device="cuda").half()
rows = [[h * (Bq // 2) + f] + [-(h * R + r) - 1 for r in range(4)] for h in range(2) for f in range(Bq // 2)]
idx = torch.tensor(rows, dtype=torch.int32, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = ops.attention(qkv, 0, 3 * C, qkv, C, 2 * C, 3 * C, refkv, C, 2 * C, 3 * C, Bq, N, N, 8, 40, idx, [0.6, .1, .1, .1, .1])
torch.cuda.synchronize()
print(float(out.float().abs().mean()))
