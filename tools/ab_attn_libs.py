"""A/B of attention-kernel builds in ONE process: every library given on the command line is loaded with ctypes, runs the
multi-source attention at the bench shape (CUDA events) and on a small shape checked against the oracle; outputs of
all libraries are compared bit for bit with the first one."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200 import _lib
from oracle import crossview_attn as cva

libs = sys.argv[1:]
ONES = os.environ.get("GCB_AB_ONES", "0") == "1"   # V heads padded to 48 columns, column 40 = 1.0 (row sums from P V)
VS = 48 if ONES else 40
res, args_t = _lib.SIGNATURES["gcb_attn_multi_fwd"]


def run(lib, qkv, refkv, idx, w, Bq, N, heads, d, reps):
    C = heads * d
    ld = 2 * C + heads * VS
    out = torch.empty((Bq, N, C), dtype=torch.float16, device="cuda")
    wv = (ctypes.c_float * len(w))(*w)
    st = torch.cuda.current_stream().cuda_stream
    es = qkv.element_size()
    call = lambda: lib.gcb_attn_multi_fwd(qkv.data_ptr(), ld, qkv.data_ptr() + C * es, qkv.data_ptr() + 2 * C * es, ld,
                                          refkv.data_ptr() + C * es, refkv.data_ptr() + 2 * C * es, ld, out.data_ptr(), C,
                                          Bq, N, N, heads, d, VS, len(w), idx.data_ptr(), wv, d ** -0.5, 1, st)
    for _ in range(2):
        rc = call()
        assert rc == 0, lib.gcb_last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        call()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def inputs(Bq, N, C, R=4, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn((Bq, N, 3 * C), device="cuda", generator=g).half()
    refkv = torch.randn((2 * R, N, 3 * C), device="cuda", generator=g).half()
    if ONES:
        def pad(t):
            v = torch.zeros(t.shape[:2] + (8, 48), dtype=t.dtype, device=t.device)
            v[..., :40] = t[..., 2 * C:].reshape(t.shape[0], t.shape[1], 8, 40)
            v[..., 40] = 1.0
            return torch.cat([t[..., :2 * C], v.reshape(t.shape[0], t.shape[1], 8 * 48)], dim=-1).contiguous()
        qkv, refkv = pad(qkv), pad(refkv)
    F = Bq // 2
    rows = [[h * F + f] + [-(h * R + r) - 1 for r in range(4)] for h in range(2) for f in range(F)]
    return qkv, refkv, torch.tensor(rows, dtype=torch.int32, device="cuda")


w = [0.6, 0.1, 0.1, 0.1, 0.1]
big = inputs(24, 4096, 320)
small = inputs(4, 256, 320, seed=1)
q, rkv = small[0].cpu().float(), small[1].cpu().float()
C = 320
vcols = lambda t: (t[..., 2 * C:].reshape(t.shape[0], t.shape[1], 8, VS)[..., :40].reshape(t.shape[0], t.shape[1], C))
ks = [q[..., C:2 * C]] + [rkv[[h * 4 + r for h in range(2) for _ in range(2)]][..., C:2 * C] for r in range(4)]
vs = [vcols(q)] + [vcols(rkv[[h * 4 + r for h in range(2) for _ in range(2)]]) for r in range(4)]
want = cva.multi_source_attention(q[..., :C], ks, vs, w, 8)
first = None
for path in libs:
    lib = ctypes.CDLL(path)
    lib.gcb_attn_multi_fwd.restype, lib.gcb_attn_multi_fwd.argtypes = res, args_t
    lib.gcb_last_error.restype = ctypes.c_char_p
    out_s, _ = run(lib, *small, w, 4, 256, 8, 40, 1)
    rel = ((out_s.cpu().float() - want).norm() / want.norm()).item()
    out_b, ms = run(lib, *big, w, 24, 4096, 8, 40, 10)
    tf = 24 * 5 * 4.0 * 4096 * 4096 * 320 / (ms / 1e3) / 1e12
    same = "ref" if first is None else str(bool(torch.equal(out_b, first[0]) and torch.equal(out_s, first[1])))
    if first is None:
        first = (out_b, out_s)
    mx = (out_s.cpu().float() - want).abs().max().item()
    print(f"{os.path.basename(path):32s} ones={int(ONES)} max|err| {mx:.2e} {ms:7.3f} ms {tf:6.1f} TFLOP/s  oracle rel {rel:.2e}  bit-identical to first: {same}", flush=True)
