"""The peer-memory exchange kernels (csrc/peer.cu) on ONE GPU: with world = 1 the push and the signal/wait kernels run
against the rank's own arena, which checks the copy, the epoch flags, the barrier, graph capture and the bounds checks on
the single-GPU test box (the cross-GPU path: tests/test_multigpu_gpu.py and bench.py's multi_gpu_check)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _handle(arena_bytes):
    from gaussctrl_b200._lib import check, lib
    h = ctypes.c_void_p()
    check(lib.gcb_handle_create(1, 0, arena_bytes, ctypes.byref(h)))
    return lib, check, h


def test_allgather_world1_copy_epochs_and_graph_capture():
    from gaussctrl_b200.parallel import _ArenaView
    lib, check, h = _handle(64 << 20)
    try:
        base = int(lib.gcb_handle_arena(h))
        off = (int(lib.gcb_handle_control_bytes()) + 255) // 256 * 256
        shape = (3, 1000, 960)
        view = torch.as_tensor(_ArenaView(base + off, shape, "<f2"), device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):                      # epochs 1..3 on the same slot
            src = torch.randn(shape, device="cuda").half()
            check(lib.gcb_allgather_ref_kv(h, off, src.data_ptr(), src.numel() * 2, 5, st))
            check(lib.gcb_peer_barrier(h, 0, st))
            torch.cuda.synchronize()
            assert torch.equal(view, src)
        # captured in a CUDA graph and replayed: the epoch counter lives in the arena, so replays keep working
        src = torch.randn(shape, device="cuda").half()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            check(lib.gcb_peer_barrier(h, 0, torch.cuda.current_stream().cuda_stream))
            check(lib.gcb_allgather_ref_kv(h, off, src.data_ptr(), src.numel() * 2, 5,
                                           torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            src.normal_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(view, src)
        err = ctypes.c_int(-1)
        check(lib.gcb_handle_error(h, ctypes.byref(err)))
        assert err.value == 0
        # bounds / alignment are rejected before anything is launched
        assert lib.gcb_allgather_ref_kv(h, off, src.data_ptr(), 64 << 20, 5, st) == -1
        assert lib.gcb_allgather_ref_kv(h, off + 8, src.data_ptr(), 1024, 5, st) == -1
        assert lib.gcb_allgather_ref_kv(h, 0, src.data_ptr(), 1024, 5, st) == -1          # would overwrite the control block
        assert lib.gcb_peer_barrier(h, 128, st) == -1
    finally:
        lib.gcb_handle_destroy(h)


def test_peer_gather_object_world1():
    """parallel.PeerKVAllGather end to end in a 1-rank process group: regions per layer name, in-place result views."""
    import os
    import socket
    import torch.distributed as dist
    from gaussctrl_b200 import parallel as par
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        gather = par.PeerKVAllGather("cuda", arena_bytes=64 << 20)
        a = torch.randn((2, 256, 960), device="cuda").half()
        b = torch.randn((2, 64, 3840), device="cuda").half()
        gather.begin_pass()
        ga, gb = gather("layerA", a), gather("layerB", b)
        torch.cuda.synchronize()
        assert torch.equal(ga, a) and torch.equal(gb, b) and ga.data_ptr() != gb.data_ptr()
        a2 = torch.randn_like(a)
        ga2 = gather("layerA", a2)
        torch.cuda.synchronize()
        assert ga2.data_ptr() == ga.data_ptr() and torch.equal(ga2, a2)     # stable address per layer: graphs can read it
        gather.check()
        # projection + exchange as one kernel (gcb_linear_allgather_fwd): equals the plain GEMM, bit for bit
        from gaussctrl_b200 import ops
        x = torch.randn((2, 256, 320), device="cuda").half()
        w = (torch.randn((960, 320), device="cuda") / 18).half()
        full, local = gather.linear_gather("layerC", x, w)
        torch.cuda.synchronize()
        want = ops.linear(x, w)
        assert full.shape == (2, 256, 960) and torch.equal(full, want) and torch.equal(local, want)
        x2 = torch.randn((1, 64, 1280), device="cuda").half()          # M = 64 < one 128-row tile
        w2 = (torch.randn((3840, 1280), device="cuda") / 36).half()
        full2, _ = gather.linear_gather("layerD", x2, w2)
        torch.cuda.synchronize()
        assert torch.equal(full2, ops.linear(x2, w2))
        assert gather.linear_gather("layerE", x, w[:100].contiguous()) is None   # Cout % 64 != 0: caller falls back
        full3, _ = gather.linear_gather("layerF", x, w, None, 320)               # world = 1: local_cols changes nothing locally
        torch.cuda.synchronize()
        assert torch.equal(full3, want)
        with pytest.raises(MemoryError):
            gather("too_big", torch.zeros((64, 4096, 960), device="cuda").half())
        gather.close()
    finally:
        dist.destroy_process_group()
