"""2-GPU test: views sharded across ranks + reference pass sharded over its CFG rows with the per-layer K/V all-gather
gives the SAME latents as the single-GPU run (every kernel is batch-invariant, the all-gather is exact).  Needs a
2-GPU box (`gpurun --gpus 2`); the same check runs inside `bench.py` at N > 1 (`extra.multi_gpu_check`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

HW = 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir, kind):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gaussctrl_b200 import parallel as par
        from gaussctrl_b200.diffusion import SD15Denoiser
        from gaussctrl_b200.engine import EditEngine
        from gaussctrl_b200.sd15_spec import synthetic_weights
        unet, cnet, _ = synthetic_weights(0, with_vae=False)
        den = SD15Denoiser(unet, cnet, f"cuda:{rank}")
        g = torch.Generator().manual_seed(7)
        V, S, guidance = 9, 2, 5.0
        lat = torch.randn((V, 4, HW, HW), generator=g).half()
        disp = torch.rand((V, 1, HW * 8, HW * 8), generator=g).repeat(1, 3, 1, 1).half()
        pos, neg = torch.randn((1, 77, 768), generator=g), torch.randn((1, 77, 768), generator=g)
        ref_idx = [0, 3, 5, 8]
        eng = EditEngine(den, use_graphs=True)
        os.environ["GCB_KV_GATHER"] = kind
        gather = par.make_kv_gather(f"cuda:{rank}", arena_bytes=1 << 30)
        assert type(gather).__name__ == ("PeerKVAllGather" if kind == "peer" else "KVAllGather"), type(gather).__name__
        ctx = {"world": world, "rank": rank, "gather": gather}
        mine = par.shard_views(V, world, rank, ref_idx)
        out = eng.edit_refs_once(lat, disp, ref_idx, pos, neg, S, guidance, view_batch=2, view_ids=mine, dist_ctx=ctx)
        full = par.gather_view_results(out[mine].contiguous(), mine, V, world)
        if rank == 0:
            for ri in ref_idx:
                full[ri] = out[ri]
            eng1 = EditEngine(den, use_graphs=True)
            want = eng1.edit_refs_once(lat, disp, ref_idx, pos, neg, S, guidance, view_batch=2)
            diff = (full.float() - want.float()).abs().max().item()
            open(os.path.join(tmpdir, "result"), "w").write(f"{diff} {ctx['gather'].bytes}")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["peer", "nccl"])
def test_two_gpu_sharded_edit_equals_single_gpu(tmp_path, kind):
    """kind = "peer": gcb_allgather_ref_kv (our NVLink peer-memory kernels, captured in the reference pass's CUDA graph);
    kind = "nccl": torch.distributed all-gather (eager)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), kind), nprocs=2, join=True)
    diff, nbytes = open(tmp_path / "result").read().split()
    assert float(diff) == 0.0, diff
    assert int(nbytes) > 0
