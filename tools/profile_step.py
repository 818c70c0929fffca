"""One denoising step of the hot path, eager (no CUDA graph), for ncu: the reference pass (R=4, CFG batch 8, K/V
recorded) followed by one view batch (c=3, CFG batch 2*c, cached reference K/V) at 512^2 (64x64 latents)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gaussctrl_b200 import ops  # noqa: E402
from gaussctrl_b200.diffusion import SD15Denoiser, cached_crossview_plan, literal_crossview_plan  # noqa: E402
from gaussctrl_b200.sd15_spec import synthetic_weights  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
unet, cnet, _ = synthetic_weights(0, with_vae=False)
den = SD15Denoiser(unet, cnet, "cuda")
g = torch.Generator().manual_seed(0)
den.set_prompts(torch.randn((2, 77, 768), generator=g))
R, c = 4, int(os.environ.get("GCB_PROFILE_VB", "12"))
rec = {}
ref_plan = literal_crossview_plan(R, "cuda", record_kv=rec)
x_ref = torch.randn((2 * R, 64, 64, 4), generator=g).half().cuda()
x_view = torch.randn((2 * c, 64, 64, 4), generator=g).half().cuda()
cond_ref = den.controlnet_cond(torch.rand((R, 512, 512, 3), generator=g).half().cuda())
cond_view = den.controlnet_cond(torch.rand((c, 512, 512, 3), generator=g).half().cuda())
t_ref = torch.full((2 * R,), 501.0, device="cuda")
t_view = torch.full((2 * c,), 501.0, device="cuda")
for it in range(reps + 1):  # first iteration = warm-up (weight packing, cudaFuncSetAttribute)
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("profiled")
        ops.GEMM_LOG = []
    e_ref = den.eps(x_ref, t_ref, torch.cat([cond_ref, cond_ref]), ref_plan)
    view_plan = cached_crossview_plan(c, R, "cuda", rec)
    e_view = den.eps(x_view, t_view, torch.cat([cond_view, cond_view]), view_plan)
torch.cuda.synchronize()
import json
os.makedirs("gpurun_out", exist_ok=True)
json.dump(ops.GEMM_LOG, open("gpurun_out/gemm_log.json", "w"))
print("launches", ops.LAUNCHES[0], "eps", float(e_view.float().abs().mean()))
