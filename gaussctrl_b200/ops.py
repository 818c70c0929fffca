"""Thin torch-tensor wrappers over the C ABI (include/gaussctrl_b200.h).

PyTorch is used here only as the owner of device memory and of the CUDA stream; every function enqueues hand-written
sm_100a kernels from libgaussctrl_b200.so on `torch.cuda.current_stream()`.  Activations are channels-last fp16:
images [B,H,W,C], tokens [B,N,C] (the same memory).  There is no CPU path: a CPU tensor raises."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (GCB_ACT_GEGLU, GCB_ACT_NONE, GCB_ACT_SILU, GCB_ATTN_AUTO, GCB_ATTN_MMA_SYNC, GCB_GEMM_MMA_SYNC,
                   GCB_GEMM_TCGEN05,
                   GCB_GEMM_TCGEN05_DIRECT, check,
                   lib)

LAUNCHES = [0]  # number of C-ABI compute calls issued (bench.py reports kernel launches from this)
ATTN_EVENTS = None  # set to a list: attention() appends (B, Nq, Nk, heads, d, n_active_sources, start_event, end_event)
GEMM_LOG = None  # set to a list to record (B, H, W, Cin, Cout, ksize, act, has_bias, has_rowvec, has_residual) per conv2d call

_GEMM_IMPL = [GCB_GEMM_TCGEN05]
_ATTN_IMPL = [GCB_ATTN_AUTO]


def set_gemm_impl(impl: int) -> None:
    _GEMM_IMPL[0] = impl


def set_attn_impl(impl: int) -> None:
    _ATTN_IMPL[0] = impl


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor], offset_elems: int = 0) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.GcbError("gaussctrl_b200 ops need CUDA tensors (there is no CPU path)")
    return t.data_ptr() + offset_elems * t.element_size()


def _f16(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float16 and t.is_contiguous(), (t.dtype, t.is_contiguous())
    return t


# ------------------------------------------------------------------------------------------------ GEMM / conv
def conv2d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], ksize: int, *, act: int = GCB_ACT_NONE,
           rowvec: Optional[torch.Tensor] = None, rowvec_off: int = 0, rowvec_ld: int = 0,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [B,H,W,Cin] fp16; w [Cout, ksize*ksize*Cin] (OHWI); returns [B,H,W,Cout] (GEGLU: Cout/2)."""
    _f16(x), _f16(w)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == ksize * ksize * Cin, (w.shape, ksize, Cin)
    co = Cout // 2 if act == GCB_ACT_GEGLU else Cout
    y = out if out is not None else torch.empty((B, H, W, co), dtype=torch.float16, device=x.device)
    if residual is not None:
        assert residual.shape == y.shape and residual.is_contiguous()
    if GEMM_LOG is not None:
        GEMM_LOG.append((B, H, W, Cin, Cout, ksize, act, bias is not None, rowvec is not None, residual is not None))
    check(lib.gcb_conv2d_nhwc_fwd(_p(x), _p(w), _p(bias), _p(rowvec, rowvec_off), rowvec_ld, _p(residual), _p(y), B, H, W,
                                  Cin, Cout, ksize, act, _GEMM_IMPL[0], _stream()))
    LAUNCHES[0] += 1
    return y


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = GCB_ACT_NONE,
           residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [..., Cin] -> [..., Cout]; w [Cout, Cin]."""
    shp = x.shape
    M = x.numel() // shp[-1]
    y = conv2d(x.reshape(1, 1, M, shp[-1]), w, bias, 1, act=act,
               residual=None if residual is None else residual.reshape(1, 1, M, -1))
    return y.reshape(*shp[:-1], y.shape[-1])


def conv2d_direct(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], ksize: int, stride: int = 1,
                  pad: Tuple[int, int] = (1, 1), act: int = GCB_ACT_NONE,
                  residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    _f16(x), _f16(w)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho = (H + pad[0] + pad[1] - ksize) // stride + 1
    Wo = (W + pad[0] + pad[1] - ksize) // stride + 1
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float16, device=x.device)
    check(lib.gcb_conv2d_direct_nhwc_fwd(_p(x), _p(w), _p(bias), _p(residual), _p(y), B, H, W, Cin, Cout, ksize, stride,
                                         pad[0], pad[1], act, _stream()))
    LAUNCHES[0] += 1
    return y


def conv3x3_s2(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], pad: Tuple[int, int] = (1, 1)):
    """Downsample2D: im2col (3x3, stride 2) + GEMM."""
    B, H, W, C = x.shape
    Ho = (H + pad[0] + pad[1] - 3) // 2 + 1
    Wo = (W + pad[0] + pad[1] - 3) // 2 + 1
    col = torch.empty((1, 1, B * Ho * Wo, 9 * C), dtype=torch.float16, device=x.device)
    check(lib.gcb_im2col3x3_s2_nhwc(_p(_f16(x)), _p(col), B, H, W, C, pad[0], pad[1], _stream()))
    LAUNCHES[0] += 1
    y = conv2d(col, w, bias, 1)
    return y.reshape(B, Ho, Wo, w.shape[0])


def im2col3x3_c4(x: torch.Tensor) -> torch.Tensor:
    """x [B,H,W,4] fp16 -> 3x3 / stride 1 / pad 1 patches [1,1,B*H*W,40] ((kh,kw,c) order, 4 zero columns)."""
    B, H, W, C = x.shape
    assert C == 4, x.shape
    col = torch.empty((1, 1, B * H * W, 40), dtype=torch.float16, device=x.device)
    check(lib.gcb_im2col3x3_c4_nhwc(_p(_f16(x)), _p(col), B, H, W, _stream()))
    LAUNCHES[0] += 1
    return col


def conv3x3_c4(x: torch.Tensor, w40: torch.Tensor, bias: Optional[torch.Tensor],
               residual: Optional[torch.Tensor] = None, col: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 convolution of a 4-channel tensor (conv_in on the latents) as a K = 40 GEMM on the tensor
    cores over its patches (`col` = im2col3x3_c4(x), shared by the UNet's and the ControlNet's conv_in).
    w40 [Cout, 40] = the OHWI weights [Cout, 36] zero-padded."""
    B, H, W, _ = x.shape
    assert w40.shape[1] == 40, w40.shape
    if col is None:
        col = im2col3x3_c4(x)
    y = conv2d(col, w40, bias, 1, residual=None if residual is None else residual.reshape(1, 1, B * H * W, -1))
    return y.reshape(B, H, W, w40.shape[0])


def geglu_perm(cout: int, device) -> torch.Tensor:
    perm = (ctypes.c_int32 * cout)()
    check(lib.gcb_geglu_pack_rows(cout, perm))
    return torch.tensor(list(perm), dtype=torch.long, device=device)


# ------------------------------------------------------------------------------------------------ norms
def groupnorm(x1: torch.Tensor, x2: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, groups: int,
              eps: float, silu: bool) -> torch.Tensor:
    """y = act(GN(cat(x1, x2, dim=-1))) over [B, H, W, C1(+C2)]."""
    _f16(x1)
    B, C1 = x1.shape[0], x1.shape[-1]
    HW = x1.numel() // (B * C1)
    C2 = 0 if x2 is None else x2.shape[-1]
    y = torch.empty(tuple(x1.shape[:-1]) + (C1 + C2,), dtype=torch.float16, device=x1.device)
    # the statistics workspace is allocated per call: inside a CUDA-graph capture it then belongs to THAT graph's pool.
    # (A workspace cached per stream was shared by every graph captured on torch's capture stream: two graphs replayed
    # concurrently raced on it, and it dangled once the graph whose pool owned it was destroyed.)
    nbytes = lib.gcb_groupnorm_workspace_bytes(B, groups)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x1.device)
    check(lib.gcb_groupnorm_nhwc_fwd(_p(x1), _p(x2), _p(gamma), _p(beta), _p(y), B, HW, C1, C2, groups, eps, int(silu),
                                     _p(ws), ws.numel(), _stream()))
    LAUNCHES[0] += 2
    return y


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _f16(x)
    C = x.shape[-1]
    y = torch.empty_like(x)
    check(lib.gcb_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), x.numel() // C, C, eps, _stream()))
    LAUNCHES[0] += 1
    return y


# ------------------------------------------------------------------------------------------------ attention
def attention(q: torch.Tensor, q_off: int, ld_q: int, kv: torch.Tensor, k_off: int, v_off: int, ld_kv: int,
              kv2: Optional[torch.Tensor], k2_off: int, v2_off: int, ld_kv2: int, B: int, Nq: int, Nk: int, heads: int,
              d: int, src_index: torch.Tensor, weights: Sequence[float], scale: Optional[float] = None,
              v_head_stride: Optional[int] = None) -> torch.Tensor:
    """Multi-source attention (see gcb_attn_multi_fwd).  q/k/v are given as (buffer, element offset, row stride) so
    they can be column slices of a fused QKV projection.  src_index: int32 [B, n_src] device tensor."""
    out = torch.empty((B, Nq, heads * d), dtype=torch.float16, device=q.device)
    n_src = len(weights)
    assert src_index.dtype == torch.int32 and src_index.numel() == B * n_src and src_index.is_cuda
    w = (ctypes.c_float * n_src)(*[float(v) for v in weights])
    sc = d ** -0.5 if scale is None else scale
    ev = None
    if ATTN_EVENTS is not None:   # in-situ timing (bench.py roofline): events on the launching stream around this kernel
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    check(lib.gcb_attn_multi_fwd(_p(q, q_off), ld_q, _p(kv, k_off), _p(kv, v_off), ld_kv, _p(kv2, k2_off),
                                 _p(kv2, v2_off), ld_kv2, _p(out), heads * d, B, Nq, Nk, heads, d,
                                 d if v_head_stride is None else v_head_stride, n_src, _p(src_index), w, sc,
                                 _ATTN_IMPL[0], _stream()))
    if ev is not None:
        ev[1].record()
        ATTN_EVENTS.append((B, Nq, Nk, heads, d, sum(1 for x in weights if float(x) != 0.0), ev[0], ev[1]))
    LAUNCHES[0] += 1
    return out


def attention_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, d: int, src_index: torch.Tensor,
                  weights: Sequence[float], scale: Optional[float] = None) -> torch.Tensor:
    """Multi-source attention on separate contiguous projections: q [B,Nq,h*d], k/v [Bk,Nk,h*d] fp16 (the layout a
    diffusers `Attention` module's to_q/to_k/to_v produce; utils.CrossViewAttnProcessor).  src_index [B,n_src] rows of k/v."""
    _f16(q), _f16(k), _f16(v)
    B, Nq, C = q.shape
    assert C == heads * d and k.shape == v.shape and k.shape[-1] == C, (q.shape, k.shape, v.shape, heads, d)
    Nk = k.shape[1]
    n_src = len(weights)
    assert src_index.dtype == torch.int32 and src_index.numel() == B * n_src and src_index.is_cuda
    out = torch.empty((B, Nq, C), dtype=torch.float16, device=q.device)
    w = (ctypes.c_float * n_src)(*[float(x) for x in weights])
    sc = d ** -0.5 if scale is None else float(scale)
    kp, vp_, ld_kv, vs = _p(k), _p(v), C, d
    if d == 40 and Nq % 128 == 0 and Nk % 128 == 0 and _ATTN_IMPL[0] != GCB_ATTN_MMA_SYNC:
        # the tcgen05 kernel's fast path at head dim 40 wants the V heads padded to 48 columns with a ones column (row
        # sums out of the P V product, part of the exponentials as packed-half polynomials: 454 -> 666 TFLOP/s at
        # B = 24, N = 4096, profiles/r3w_time_attn.txt).  K and V share one row stride in the ABI: one [k | v padded]
        # buffer, a copy of K and V (1 % of the attention's own traffic at N = 4096)
        Bk, ldp = k.shape[0], C + heads * 48
        kv = torch.zeros((Bk, Nk, ldp), dtype=torch.float16, device=k.device)
        kv[..., :C] = k
        vv = kv[..., C:].view(Bk, Nk, heads, 48)
        vv[..., :d] = v.view(Bk, Nk, heads, d)
        vv[..., d] = 1.0
        kp, vp_, ld_kv, vs = _p(kv), _p(kv, C), ldp, 48
        LAUNCHES[0] += 4
    check(lib.gcb_attn_multi_fwd(_p(q), C, kp, vp_, ld_kv, None, None, 0, _p(out), C, B, Nq, Nk, heads, d, vs, n_src,
                                 _p(src_index), w, sc, _ATTN_IMPL[0], _stream()))
    LAUNCHES[0] += 1
    return out


def softmax_rows(x: torch.Tensor, scale: float) -> torch.Tensor:
    y = torch.empty_like(_f16(x))
    cols = x.shape[-1]
    check(lib.gcb_softmax_rows_fwd(_p(x), _p(y), x.numel() // cols, cols, scale, _stream()))
    LAUNCHES[0] += 1
    return y


# ------------------------------------------------------------------------------------------------ elementwise
def silu(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(_f16(x))
    check(lib.gcb_silu_fwd(_p(x), _p(y), x.numel(), _stream()))
    LAUNCHES[0] += 1
    return y


def add(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0, beta: float = 1.0) -> torch.Tensor:
    assert a.shape == b.shape
    y = torch.empty_like(_f16(a))
    check(lib.gcb_add_fwd(_p(a), _p(_f16(b)), _p(y), a.numel(), alpha, beta, _stream()))
    LAUNCHES[0] += 1
    return y


def geglu(x: torch.Tensor) -> torch.Tensor:
    C = x.shape[-1] // 2
    y = torch.empty(tuple(x.shape[:-1]) + (C,), dtype=torch.float16, device=x.device)
    check(lib.gcb_geglu_fwd(_p(_f16(x)), _p(y), x.numel() // (2 * C), C, _stream()))
    LAUNCHES[0] += 1
    return y


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    B, H, W, C = x.shape
    y = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float16, device=x.device)
    check(lib.gcb_upsample_nearest2x_nhwc(_p(_f16(x)), _p(y), B, H, W, C, _stream()))
    LAUNCHES[0] += 1
    return y


def timestep_embedding(t_dev: torch.Tensor, dim: int) -> torch.Tensor:
    assert t_dev.dtype == torch.float32 and t_dev.is_cuda
    B = t_dev.numel()
    y = torch.empty((B, dim), dtype=torch.float16, device=t_dev.device)
    check(lib.gcb_timestep_embedding(_p(t_dev), B, dim, _p(y), _stream()))
    LAUNCHES[0] += 1
    return y


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    B, C, H, W = x.shape
    y = torch.empty((B, H, W, C), dtype=torch.float16, device=x.device)
    check(lib.gcb_nchw_to_nhwc_f16(_p(_f16(x)), _p(y), B, C, H, W, _stream()))
    LAUNCHES[0] += 1
    return y


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    B, H, W, C = x.shape
    y = torch.empty((B, C, H, W), dtype=torch.float16, device=x.device)
    check(lib.gcb_nhwc_to_nchw_f16(_p(_f16(x)), _p(y), B, C, H, W, _stream()))
    LAUNCHES[0] += 1
    return y


def transpose(x: torch.Tensor) -> torch.Tensor:
    """[batch, rows, cols] -> [batch, cols, rows]."""
    b, r, c = x.shape
    y = torch.empty((b, c, r), dtype=torch.float16, device=x.device)
    check(lib.gcb_transpose_f16(_p(_f16(x)), _p(y), b, r, c, _stream()))
    LAUNCHES[0] += 1
    return y


def cfg_ddim_step(eps_uncond: torch.Tensor, eps_cond: Optional[torch.Tensor], x: torch.Tensor, guidance: float,
                  coef_dev: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """coef_dev: device fp32[4] = sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)."""
    y = out if out is not None else torch.empty_like(_f16(x))
    assert eps_uncond.numel() == x.numel() and coef_dev.dtype == torch.float32
    check(lib.gcb_cfg_ddim_step(_p(eps_uncond), _p(eps_cond), _p(x), _p(y), x.numel(), guidance, _p(coef_dev),
                                _stream()))
    LAUNCHES[0] += 1
    return y


def postprocess_composite(img: torch.Tensor, mask: Optional[torch.Tensor], unedited: Optional[torch.Tensor]):
    """img [B,H,W,3] fp16 (decoder output) -> [B,H,W,3] fp32 = clamp(img/2+0.5) (* mask + unedited * (1-mask))."""
    B, H, W, _ = img.shape
    out = torch.empty((B, H, W, 3), dtype=torch.float32, device=img.device)
    check(lib.gcb_postprocess_composite(_p(_f16(img)), _p(mask), _p(unedited), _p(out), B, H, W, _stream()))
    LAUNCHES[0] += 1
    return out


def depth_to_disparity(depth: torch.Tensor, round_f16_first: bool) -> torch.Tensor:
    """depth fp32 [B,H,W] -> disparity fp16 [B,H,W,3]."""
    assert depth.dtype == torch.float32 and depth.is_contiguous()
    B, H, W = depth.shape
    ws = torch.empty(B, dtype=torch.float32, device=depth.device)
    out = torch.empty((B, H, W, 3), dtype=torch.float16, device=depth.device)
    check(lib.gcb_depth_to_disparity(_p(depth), _p(out), _p(ws), B, H * W, int(round_f16_first), _stream()))
    LAUNCHES[0] += 2
    return out
