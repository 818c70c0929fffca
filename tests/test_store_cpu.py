"""CPU tests of the inter-stage store: the reference's on-disk layout of the stage-A products
(gc_dataparser_ns.py:408-420, gc_dataset.py:35-66,129-162, gc_render.py:217-221) round-trips `train_data`, and
diffusers-layout safetensors checkpoints load into the sd15_spec parameter tables (gc_pipeline.py:97-102)."""
import os

import numpy as np
import pytest
import torch


def _train_data(V, H, W, with_mask=True):
    g = torch.Generator().manual_seed(0)
    td = []
    for i in range(V):
        e = {"image_idx": i,
             "unedited_image": torch.rand((H, W, 3), generator=g).to(torch.float16),
             "depth_image": (torch.rand((1, H, W), generator=g) * 5 + 0.1).numpy(),
             "z_0_image": torch.randn((1, 4, H // 8, W // 8), generator=g).numpy()}
        if with_mask:
            e["mask_image"] = (torch.rand((H, W), generator=g) > 0.5).numpy() * 1
        td.append(e)
    return td


def test_layout_and_round_trip(tmp_path):
    from gaussctrl_b200 import store
    root = str(tmp_path)
    assert store.available(root) == [] and not store.has_stage_a(root)
    td = _train_data(3, 64, 48)
    store.save_train_data(root, td)
    # the reference's names: 1-based frame index, these four folders
    for folder, ext in (("depth_npy", "npy"), ("z_0", "npy"), ("mask_npy", "npy"), ("unedited", "jpg")):
        assert sorted(os.listdir(os.path.join(root, folder))) == [f"frame_{i:05d}.{ext}" for i in (1, 2, 3)]
    # depth is stored as the model output [H,W,1] f32 (gc_render.py:221) ...
    raw = np.load(os.path.join(root, "depth_npy", "frame_00002.npy"))
    assert raw.shape == (64, 48, 1) and raw.dtype == np.float32
    # ... and read back as the reference's dataset does: np.load(p)[:, :, 0][None]
    assert store.has_stage_a(root)
    back = store.load_train_data(root, 3)
    for a, b in zip(td, back):
        assert b["depth_image"].shape == (1, 64, 48) and np.array_equal(a["depth_image"], b["depth_image"])
        assert b["z_0_image"].shape == (1, 4, 8, 6) and np.array_equal(a["z_0_image"], b["z_0_image"])
        assert np.array_equal(a["mask_image"], b["mask_image"])
        u = b["unedited_image"]
        assert u.dtype == torch.float32 and u.shape == (64, 48, 3) and float(u.min()) >= 0 and float(u.max()) <= 1
        # JPEG is lossy on noise; the 8-bit quantisation itself is exact on a flat image (below)
        assert (u - a["unedited_image"].float()).abs().mean().item() < 0.2
    assert "mask_image" not in store.load_view(root, 0, load_mask=False)


def test_unedited_jpeg_quantisation(tmp_path):
    from gaussctrl_b200 import store
    img = torch.full((32, 32, 3), 0.5)
    img[:, :, 1] = 0.25
    store.save_view(str(tmp_path), 4, {"unedited_image": img})
    assert os.path.isfile(os.path.join(str(tmp_path), "unedited", "frame_00005.jpg"))
    back = store.load_view(str(tmp_path), 4)["unedited_image"]
    assert (back - img).abs().max().item() <= 2.0 / 255.0


def test_partial_folders(tmp_path):
    from gaussctrl_b200 import store
    td = _train_data(2, 16, 16, with_mask=False)
    for e in td:
        del e["unedited_image"]
    store.save_train_data(str(tmp_path), td)
    assert sorted(store.available(str(tmp_path))) == ["depth_image", "z_0_image"]
    back = store.load_train_data(str(tmp_path), 2, [{"image_idx": 0, "image": "keep"}, {"image_idx": 1}])
    assert back[0]["image"] == "keep" and set(back[1]) == {"image_idx", "depth_image", "z_0_image"}


def test_checkpoint_round_trip_and_validation(tmp_path):
    """Small stand-in for the 4 GB checkpoint: the VAE table only (same code path for unet/controlnet)."""
    from safetensors.torch import save_file
    from gaussctrl_b200 import checkpoint as ck, sd15_spec as sp
    shapes = sp.vae_shapes()
    sd = sp.random_state_dict(shapes, seed=5)
    folder = os.path.join(str(tmp_path), "vae")
    os.makedirs(folder)
    # write with the legacy SD1.x attention names, conv-shaped, as the published VAE checkpoints have them
    legacy = {}
    inv = {v: k for k, v in ck._VAE_ATTN_RENAME.items()}
    for k, v in sd.items():
        if ".attentions." in k:
            for new, old in inv.items():
                if f".{new}." in k:
                    k = k.replace(f".{new}.", f".{old}.")
                    if k.endswith(".weight"):
                        v = v[:, :, None, None]
                    break
        legacy[k] = v.half().contiguous()
    save_file(legacy, os.path.join(folder, "diffusion_pytorch_model.safetensors"))
    got = ck.load_component(folder, shapes, "vae", vae=True)
    assert set(got) == set(shapes)
    for k in shapes:
        assert got[k].dtype == torch.float32 and torch.equal(got[k], sd[k].half().float())
    # validation: a missing tensor and a wrong shape are both refused
    bad = dict(legacy)
    bad.pop("quant_conv.weight")
    save_file(bad, os.path.join(folder, "diffusion_pytorch_model.safetensors"))
    with pytest.raises(ValueError, match="missing"):
        ck.load_component(folder, shapes, "vae", vae=True)
    bad = dict(legacy)
    bad["quant_conv.weight"] = torch.zeros(8, 4, 1, 1).half()
    save_file(bad, os.path.join(folder, "diffusion_pytorch_model.safetensors"))
    with pytest.raises(ValueError, match="shape"):
        ck.load_component(folder, shapes, "vae", vae=True)
    with pytest.raises(FileNotFoundError):
        ck.load_component(os.path.join(str(tmp_path), "unet"), sp.unet_shapes(), "unet")
