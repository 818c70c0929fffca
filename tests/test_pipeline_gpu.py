"""End-to-end parity of the plugin surface: GaussCtrlPipeline.render_reverse() + edit_images() (public API, host
`train_data` in / out) against the oracle's composition of the reference's two loops
(gc_pipeline.py:122-157 and :159-237: rasterise -> image2latent -> depth2disparity_torch -> DDIM inversion ->
per-chunk cross-view edit with refs recomputed -> VAE decode -> mask composite) on a small seeded scene."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 256          # 32x32 latents
V, R, CHUNK, S, GUIDANCE = 6, 4, 2, 2, 5.0


def _scene(n, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(means=torch.rand((n, 3), generator=g) * 1.6 - 0.8, scales=torch.randn((n, 3), generator=g) * 0.3 + math.log(0.06),
                quats=torch.randn((n, 4), generator=g), opacities=torch.rand((n, 1), generator=g) * 6 - 1,
                features_dc=torch.randn((n, 3), generator=g) * 0.5, features_rest=torch.randn((n, 15, 3), generator=g) * 0.05)


def _c2w(i, n, radius=2.4):
    az = 2 * math.pi * i / n
    eye = torch.tensor([radius * math.cos(az), radius * math.sin(az), 0.5])
    fwd = -eye / eye.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    m = torch.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
    return m


def test_render_reverse_and_edit_images_match_oracle():
    from oracle import gsplat_ref as gr, pipeline as opipe, sd15
    from gaussctrl_b200._compat import Cameras
    from gaussctrl_b200.gc_model import GaussCtrlModel, GaussCtrlModelConfig
    from gaussctrl_b200.gc_pipeline import (GaussCtrlPipeline, GaussCtrlPipelineConfig, SimpleDataManager,
                                            select_ref_indices, synthetic_prompt_embeds)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    unet, cnet, vae = sd15.seeded_models(seed=0, with_vae=True)
    weights = (unet.state_dict(), cnet.state_dict(), vae.state_dict())
    P = _scene(3000, seed=3)
    fx = fy = 1.05 * W
    cx, cy = W / 2 + 1.5, H / 2 - 2.0
    c2ws = torch.stack([_c2w(i, V) for i in range(V)])
    bg = torch.tensor([0.05, 0.1, 0.15])
    yy, xx = np.mgrid[0:H, 0:W]
    disc = ((yy - H / 2) ** 2 + (xx - W / 2) ** 2 < (0.3 * W) ** 2)

    # ---------------- product: public API
    model = GaussCtrlModel(GaussCtrlModelConfig(), num_points=3000)
    with torch.no_grad():
        for k, v in P.items():
            getattr(model, k).data = v.clone()
    model = model.cuda()
    model.background_color = bg.clone()
    dm = SimpleDataManager(Cameras(c2ws[:, :3], fx, fy, cx, cy, W, H))
    cfg = GaussCtrlPipelineConfig(edit_prompt="a", reverse_prompt="b", langsam_obj="thing", guidance_scale=GUIDANCE,
                                  num_inference_steps=S, chunk_size=CHUNK, ref_view_num=R)
    pipe = GaussCtrlPipeline(cfg, "cuda", datamanager=dm, model=model, weights=weights,
                             mask_fn=lambda rgb, obj: disc)
    assert pipe.ref_indices == select_ref_indices(V, R)
    pipe.render_reverse()
    pipe.edit_images()
    td = dm.train_data
    assert td[0]["z_0_image"].shape == (1, 4, H // 8, W // 8) and td[0]["z_0_image"].dtype == np.float32
    assert td[0]["depth_image"].shape == (1, H, W) and td[0]["unedited_image"].dtype == torch.float16
    assert td[0]["image"].shape == (H, W, 3) and td[0]["image"].dtype == torch.float32 and not td[0]["image"].is_cuda

    # ---------------- oracle: the reference's loops, fp32
    unet, cnet, vae = unet.cuda(), cnet.cuda(), vae.cuda()
    tables = sd15.DDIMTables()
    emb_rev = synthetic_prompt_embeds([pipe.positive_reverse_prompt]).half().float().cuda()
    emb = synthetic_prompt_embeds([pipe.negative_prompts, pipe.positive_prompt]).half().float().cuda()
    zT, disp_np, rgb_un = [], [], []
    for i in range(V):
        out = gr.get_outputs(P, c2ws[i], fx, fy, cx, cy, H, W, 3, bg)
        rgb16, depth16 = out["rgb"].half(), out["depth"].half()                       # gc_pipeline.py:132-133
        z0 = opipe.image2latent(vae, rgb16.float().cuda())
        disparity = opipe.depth2disparity_torch(depth16[:, :, 0][None]).float().cuda()  # fp16 arithmetic, :139
        z = opipe.invert_view(unet, cnet, tables, z0, disparity, emb_rev, S)
        zT.append(z.cpu().numpy().astype(np.float32))                                  # update_datasets :268-274
        disp_np.append(opipe.depth2disparity(depth16.permute(2, 0, 1).float().numpy()))
        rgb_un.append(rgb16)
    # stage-A products agree
    for i in range(V):
        rel = np.linalg.norm(td[i]["z_0_image"] - zT[i]) / np.linalg.norm(zT[i])
        assert rel < 2e-2, ("z_T", i, rel)  # fp16 VAE encoder + 2 inversion steps vs fp32
        assert (td[i]["unedited_image"].float() - rgb_un[i].float()).abs().max().item() < 2e-3
    ref_idx = pipe.ref_indices
    worst = 0.0
    for c0 in range(0, V, CHUNK):
        ids = list(range(c0, min(V, c0 + CHUNK)))
        sel = ref_idx + ids
        lat = torch.from_numpy(np.concatenate([zT[j] for j in sel])).half().float().cuda()
        dsp = torch.from_numpy(np.concatenate([disp_np[j] for j in sel])).half().float().cuda()
        imgs = opipe.edit_chunk(unet, cnet, vae, tables, lat, dsp, emb[1:2], emb[0:1], S, GUIDANCE, R)
        for k, j in enumerate(ids):
            want = opipe.composite_mask(imgs[k].cpu(), rgb_un[j].float(), disc.astype(np.float32))
            err = (td[j]["image"] - want).abs()
            worst = max(worst, err.mean().item())
            assert err.mean().item() < 1e-2 and err.max().item() < 0.15, (j, err.mean().item(), err.max().item())
            outside = torch.from_numpy(~disc)
            # the mask composite keeps this run's own un-edited render bit for bit outside the mask (gc_pipeline.py:227-232)
            assert torch.equal(td[j]["image"][outside], td[j]["unedited_image"].float()[outside])
    print("worst mean abs image error", worst)
