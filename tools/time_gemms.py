"""Per-shape timing of every GEMM / implicit-GEMM conv of one denoising step (reference pass, CFG batch 8, + one view
batch, CFG batch 2*vb) with CUDA events, two implementations side by side (GCB_TIME_IMPLS="old,new" impl ids of
include/gaussctrl_b200.h; default "4,3": one tile per CTA vs the persistent schedule; "2,0": round-1 per-thread row
stores vs the product default).  Prints a table sorted by time and writes gpurun_out/gemm_table.json."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import ops
from gaussctrl_b200._lib import GCB_ACT_GEGLU, GCB_GEMM_TCGEN05, GCB_GEMM_TCGEN05_DIRECT
from gaussctrl_b200.diffusion import SD15Denoiser, cached_crossview_plan, literal_crossview_plan
from gaussctrl_b200.sd15_spec import synthetic_weights

vb = int(os.environ.get("GCB_PROFILE_VB", "12"))
IMPL_OLD, IMPL_NEW = (int(v) for v in os.environ.get("GCB_TIME_IMPLS", "4,3").split(","))
unet, cnet, _ = synthetic_weights(0, with_vae=False)
den = SD15Denoiser(unet, cnet, "cuda")
g = torch.Generator().manual_seed(0)
den.set_prompts(torch.randn((2, 77, 768), generator=g))
R = 4
rec = {}
ref_plan = literal_crossview_plan(R, "cuda", record_kv=rec)
x_ref = torch.randn((2 * R, 64, 64, 4), generator=g).half().cuda()
x_view = torch.randn((2 * vb, 64, 64, 4), generator=g).half().cuda()
cond_ref = den.controlnet_cond(torch.rand((R, 512, 512, 3), generator=g).half().cuda())
cond_view = den.controlnet_cond(torch.rand((vb, 512, 512, 3), generator=g).half().cuda())
t_ref, t_view = torch.full((2 * R,), 501.0, device="cuda"), torch.full((2 * vb,), 501.0, device="cuda")
logs = {}
for it in range(2):
    ops.GEMM_LOG = []
    den.eps(x_ref, t_ref, torch.cat([cond_ref, cond_ref]), ref_plan)
    logs["ref"] = ops.GEMM_LOG
    ops.GEMM_LOG = []
    den.eps(x_view, t_view, torch.cat([cond_view, cond_view]), cached_crossview_plan(vb, R, "cuda", rec))
    logs["view"] = ops.GEMM_LOG
ops.GEMM_LOG = None
torch.cuda.synchronize()
# weights: one view batch per 12 views; the reference pass once per 36 non-reference views (3 view batches)
counts = collections.Counter()
for s in logs["view"]:
    counts[tuple(s)] += 3
for s in logs["ref"]:
    counts[tuple(s)] += 1


def time_shape(shape, impl, reps=10):
    B, H, W, Cin, Cout, k, act, hb, hr, hres = shape
    torch.manual_seed(0)  # same operands for both epilogues
    x = torch.randn((B, H, W, Cin), device="cuda").half()
    w = (torch.randn((Cout, k * k * Cin), device="cuda") / (k * k * Cin) ** 0.5).half()
    bias = torch.randn((Cout,), device="cuda").half() if hb else None
    co = Cout // 2 if act == GCB_ACT_GEGLU else Cout
    res = torch.randn((B, H, W, co), device="cuda").half() if hres else None
    rv = torch.randn((B, Cout), device="cuda").half() if hr else None
    ops.set_gemm_impl(impl)
    call = lambda: ops.conv2d(x, w, bias, k, act=act, rowvec=rv, rowvec_ld=Cout if hr else 0, residual=res)
    for _ in range(2):
        y = call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        call()
    e1.record(); torch.cuda.synchronize()
    ops.set_gemm_impl(GCB_GEMM_TCGEN05)
    return e0.elapsed_time(e1) / reps * 1e3, y   # us


rows = []
for shape, n in counts.items():
    B, H, W, Cin, Cout, k, act, hb, hr, hres = shape
    M, K = B * H * W, k * k * Cin
    flops = 2.0 * M * K * Cout
    co = Cout // 2 if act == GCB_ACT_GEGLU else Cout
    byts = 2.0 * (M * Cin + Cout * K + M * co * (2 if hres else 1))
    us_new, y_new = time_shape(shape, IMPL_NEW)
    us_old, y_old = time_shape(shape, IMPL_OLD)
    same = bool(torch.equal(y_new, y_old)) if y_new.shape == y_old.shape else False
    rows.append(dict(shape=list(shape), count=n, us_tma=us_new, us_direct=us_old, tflops_tma=flops / us_new / 1e6,
                     tflops_direct=flops / us_old / 1e6, hbm_floor_us=byts / 6.5e6, total_us_tma=n * us_new,
                     total_us_direct=n * us_old, bit_identical=same))
rows.sort(key=lambda r: -r["total_us_direct"])
tot_new, tot_old = sum(r["total_us_tma"] for r in rows), sum(r["total_us_direct"] for r in rows)
print(f"{'B':>3} {'HxW':>7} {'Cin':>5} {'Cout':>5} k act b/rv/res  n | us old -> new     | TF/s old -> new    | HBM floor us | share(old) same")
for r in rows:
    B, H, W, Cin, Cout, k, act, hb, hr, hres = r["shape"]
    print(f"{B:>3} {H:>3}x{W:<3} {Cin:>5} {Cout:>5} {k} {act}   {int(hb)}/{int(hr)}/{int(hres)}   {r['count']:>3} | "
          f"{r['us_direct']:8.1f} -> {r['us_tma']:8.1f} | {r['tflops_direct']:6.0f} -> {r['tflops_tma']:6.0f} | "
          f"{r['hbm_floor_us']:8.1f} | {r['total_us_direct'] / tot_old:6.1%} {r['bit_identical']}")
best = sum(min(r["total_us_tma"], r["total_us_direct"]) for r in rows)
print(f"impl {IMPL_OLD} -> impl {IMPL_NEW}; per-shape best of both: {best / 1e3:.2f} ms")
print(f"GEMM time per 36 views x 1 DDIM step: old {tot_old / 1e3:.2f} ms -> new {tot_new / 1e3:.2f} ms; "
      f"all outputs bit-identical: {all(r['bit_identical'] for r in rows)}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(view_batch=vb, impl_old=IMPL_OLD, impl_new=IMPL_NEW, rows=rows, total_ms_direct=tot_old / 1e3, total_ms_tma=tot_new / 1e3),
          open(os.environ.get("GCB_GEMM_TABLE_OUT", "gpurun_out/gemm_table.json"), "w"), indent=1)
