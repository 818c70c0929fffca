"""One GEMM layer shape alone for `ncu --set full`: default = the GEGLU projection 320 -> 2560 of the 64x64 level at the
production batch (M = 80 x 4096), the largest single GEMM of a denoising step.  argv: B H W Cin Cout k act residual."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import ops

a = [int(v) for v in sys.argv[1:]] or [1, 1, 327680, 320, 2560, 1, 2, 0]
B, H, W, Cin, Cout, k, act, has_res = a
torch.manual_seed(0)
x = torch.randn((B, H, W, Cin), device="cuda").half()
w = (torch.randn((Cout, k * k * Cin), device="cuda") / (k * k * Cin) ** 0.5).half()
bias = torch.randn((Cout,), device="cuda").half()
co = Cout // 2 if act == 2 else Cout
res = torch.randn((B, H, W, co), device="cuda").half() if has_res else None
for _ in range(3):
    y = ops.conv2d(x, w, bias, k, act=act, residual=res)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
