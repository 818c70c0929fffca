"""A few fine-tune iterations (training render -> fused L1+SSIM -> backward -> FusedAdam) of the synthetic 1M-Gaussian
scene at 512x512, the last one inside the NVTX range "profiled" for the ncu launch list; prints CUDA-event timings of
the loss call and the Adam launch alone (their roofline numerators: DESIGN.md §5)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_scene, orbit_c2w
from gaussctrl_b200._compat import Cameras
from gaussctrl_b200.finetune import FineTuner, l1_ssim_loss
from gaussctrl_b200.gc_model import GaussCtrlModel, GaussCtrlModelConfig
from gaussctrl_b200.gc_pipeline import SimpleDataManager

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
V = 4
scene = synthetic_scene(n)
model = GaussCtrlModel(GaussCtrlModelConfig(), num_points=n)
with torch.no_grad():
    for k, v in scene.items():
        getattr(model, k).data = v
model = model.cuda()
cams = Cameras(torch.stack([orbit_c2w(i, 40) for i in range(V)]), 539.05, 538.17, 258.74, 239.35, 512, 512)
g = torch.Generator().manual_seed(0)
dm = SimpleDataManager(cams, [{"image_idx": i, "image": torch.rand((512, 512, 3), generator=g)} for i in range(V)])
tuner = FineTuner(model, dm)
random.seed(0)
# autograd's backward normally runs on its own worker thread, outside this thread's NVTX range: keep it here so that
# `ncu --nvtx-include profiled/` also lists the backward kernels
torch.autograd.set_multithreading_enabled(False)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(4):
    if it == 3:
        torch.cuda.synchronize(); torch.cuda.nvtx.range_push("profiled")
    loss, _, _ = tuner.train_iteration(30000 + it)
    if it == 3:
        torch.cuda.synchronize(); torch.cuda.nvtx.range_pop()
print("loss", float(loss))
# the two new kernels alone, CUDA events, 50 back-to-back calls
pred, gt = torch.rand((512, 512, 3), device="cuda", generator=None), torch.rand((512, 512, 3), device="cuda")
for name, fn in (("l1_ssim_fwd_bwd", lambda: l1_ssim_loss(pred, gt)), ("adam_step", lambda: tuner.optimizer.step())):
    for _ in range(5):
        fn()
    e0, e1 = ev(), ev()
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    if name == "adam_step":
        byts = 28.0 * sum(p.numel() for grp in tuner.optimizer.param_groups for p in grp["params"] if p.grad is not None)
    else:
        byts = 512 * 512 * 3 * 4 * 3.0   # read pred, gt, write v_pred (workspace passes stay in L2)
    print(f"{name}: {ms*1e3:.1f} us per call, algorithmic {byts/1e6:.1f} MB -> {byts/ms/1e6:.0f} GB/s")
