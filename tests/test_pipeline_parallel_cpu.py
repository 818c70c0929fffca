"""world_size-2 and -3 gloo tests (CPU) of the PIPELINE-level multi-GPU host logic: render_reverse() deals the views
round-robin, edit_images() deals the non-reference views and the reference decodes, and after each call EVERY rank's
train_data is complete and identical to the single-process result (the fine-tune samples any view on any
rank).  The device compute (rasteriser, VAE, denoiser) is replaced by deterministic CPU stand-ins - the kernels
themselves are covered by the -m gpu tests; this checks sharding, index remapping and the gathers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

V, R, HW = 11, 4, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Model:
    """get_outputs_for_camera stand-in: a render that encodes the camera's x translation (= view id)."""

    def get_outputs_for_camera(self, cam):
        v = float(cam.camera_to_worlds[0, 0, 3])
        rgb = torch.full((HW, HW, 3), v / 64.0)
        return {"rgb": rgb, "depth": torch.full((HW, HW, 1), 1.0 + v), "accumulation": torch.ones(HW, HW, 1)}


def _build(world, set_attr=setattr):
    from gaussctrl_b200 import gc_pipeline as gp, ops
    from gaussctrl_b200._compat import Cameras
    c2w = torch.eye(4)[:3].repeat(V, 1, 1)
    c2w[:, 0, 3] = torch.arange(V, dtype=torch.float32)
    dm = gp.SimpleDataManager(Cameras(c2w, 20.0, 20.0, 8.0, 8.0, HW, HW))
    cfg = gp.GaussCtrlPipelineConfig(edit_prompt="a", reverse_prompt="b", langsam_obj="thing", ref_view_num=R,
                                     num_inference_steps=2, chunk_size=3)
    tiny = {f"down_blocks.{i}.resnets.0.conv1.weight": torch.zeros(8, 4, 3, 3) for i in range(4)}
    pipe = gp.GaussCtrlPipeline(cfg, "cpu", world_size=world, datamanager=dm, model=_Model(), weights=(tiny, tiny, None),
                                mask_fn=lambda rgb, text: (np.arange(HW * HW).reshape(HW, HW) % 3 == 0))
    # ---- CPU stand-ins for the device stages
    pipe.image2latent_batch = lambda images: images[:, :2, :2, :].permute(0, 3, 1, 2).repeat(1, 2, 1, 1)[:, :4].contiguous()
    set_attr(ops, "depth_to_disparity", lambda depth, f16: (1.0 / depth)[..., None].repeat(1, 1, 1, 3).half())
    set_attr(ops, "nhwc_to_nchw", lambda x: x.permute(0, 3, 1, 2).contiguous())

    class _Engine:
        def invert(self, z0, disparity, emb, S, batch=8):
            return (z0.float() + disparity[:, :1, :2, :2].float()).half()

        def edit_refs_once(self, latents, disparity, ref_indices, pos, neg, S, g, view_batch=4, view_ids=None,
                           dist_ctx=None, ref_frames=(0, 1, 2, 3), stop_after=None):
            # every view's result depends on its own latent and on ALL reference latents (as the real pass does)
            ref_sum = latents[list(ref_indices)].float().sum(dim=0, keepdim=True)
            out = torch.zeros_like(latents, dtype=torch.float32)
            ids = list(range(latents.shape[0])) if view_ids is None else list(view_ids)
            for v in set(ids) | set(ref_indices):
                out[v] = latents[v].float() * 2 + ref_sum[0] + disparity[v, :1, :2, :2].float()
            assert dist_ctx is None or (dist_ctx["world"] > 1 and dist_ctx["gather"] is not None)
            return out.half()

    class _Vae:
        def decode_latents(self, lat, masks=None, uned=None, batch=4):
            img = lat.float().mean(dim=(1, 2, 3))[:, None, None, None].expand(-1, HW, HW, 3)
            if masks is not None:
                m = masks.float()[..., None]
                img = img * m + uned.float() * (1 - m)
            return img.contiguous()

    pipe.engine, pipe.vae = _Engine(), _Vae()
    return pipe, dm


def _snapshot(dm):
    keys = ("unedited_image", "depth_image", "z_0_image", "mask_image", "image")
    return [{k: torch.as_tensor(np.asarray(d[k])).double() for k in keys} for d in dm.train_data]


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaussctrl_b200 import parallel as par
        par.make_kv_gather = lambda dev, group=None: par.KVAllGather(group)
        pipe, dm = _build(world)
        pipe.render_reverse()
        assert all("z_0_image" in d and "mask_image" in d for d in dm.train_data), "stage-A products missing on a rank"
        pipe.edit_images()
        assert all("image" in d for d in dm.train_data), "edited images missing on a rank"
        # upload accounting: a rank only uploads its own views + the references
        n_need = len(set(par.shard_views(V, world, rank, pipe.ref_indices)) | set(pipe.ref_indices))
        assert pipe.h2d_bytes >= n_need * (4 * 2 * 2 + HW * HW) * 4 and n_need < V
        torch.save(_snapshot(dm), os.path.join(tmpdir, f"rank{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_every_rank_ends_with_complete_train_data(tmp_path, world, monkeypatch):
    pipe, dm = _build(1, monkeypatch.setattr)
    pipe.render_reverse()
    pipe.edit_images()
    want = _snapshot(dm)
    assert want[0]["image"].shape == (HW, HW, 3) and want[0]["z_0_image"].shape == (1, 4, 2, 2)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        got = torch.load(os.path.join(tmp_path, f"rank{rank}.pt"))
        for v in range(V):
            for k in want[v]:
                assert torch.equal(got[v][k], want[v][k]), (rank, v, k)


def test_reference_decodes_are_dealt_to_the_ranks_with_fewer_views():
    from gaussctrl_b200 import parallel as par
    refs = [4, 11, 29, 31]
    for world in (2, 4, 8):
        owners = [par.ref_decode_owner(4, world, r) for r in range(world)]
        assert sorted(i for o in owners for i in o) == [0, 1, 2, 3]
        loads = [len(par.shard_views(40, world, r, refs)) + len(owners[r]) for r in range(world)]
        assert max(loads) - min(loads) <= 1, (world, loads)
