"""Runs in a SUBPROCESS with tests/fake_nerfstudio on sys.path (so gaussctrl_b200._compat takes its real-nerfstudio
branch): drives the plugin exactly as `ns-train gaussctrl` would - entry point -> MethodSpecification -> TrainerConfig.setup
-> Trainer.setup -> VanillaPipeline.__init__ contract -> GaussCtrlPipeline - and checks get_outputs' side effects."""
import importlib
import sys

import torch


def main():
    from gaussctrl_b200 import _compat
    assert _compat.HAVE_NERFSTUDIO, "the fake nerfstudio tree was not picked up"
    import nerfstudio.pipelines.base_pipeline as nbp
    import gaussctrl_b200.gc_pipeline as gp
    assert issubclass(gp.GaussCtrlPipeline, nbp.VanillaPipeline) and issubclass(gp.GaussCtrlPipelineConfig,
                                                                                 nbp.VanillaPipelineConfig)

    # ---- the entry point string of pyproject.toml resolves to a MethodSpecification
    import tomllib
    ep = tomllib.load(open("pyproject.toml", "rb"))["project"]["entry-points"]["nerfstudio.method_configs"]["gaussctrl"]
    mod, attr = ep.split(":")
    assert (mod, attr) == ("gaussctrl.gc_config", "gaussctrl_method")     # reference pyproject.toml:38-39
    spec = getattr(importlib.import_module(mod), attr)
    cfg = spec.config
    assert cfg.method_name == "gaussctrl" and cfg.max_num_iterations == 1000 and cfg.steps_per_save == 250
    assert cfg.steps_per_eval_image == 100 and cfg.mixed_precision is False
    assert cfg.gradient_accumulation_steps == {"camera_opt": 100}
    assert sorted(cfg.optimizers) == ["camera_opt", "features_dc", "features_rest", "opacity", "rotation", "scaling", "xyz"]
    assert cfg.optimizers["xyz"]["optimizer"].lr == 1.6e-4 and cfg.optimizers["xyz"]["scheduler"].lr_final == 1.6e-6
    assert cfg.optimizers["features_rest"]["optimizer"].lr == 0.0025 / 20 and cfg.optimizers["opacity"]["optimizer"].lr == 0.05
    assert all(o["optimizer"].eps == 1e-15 for o in cfg.optimizers.values())
    assert type(cfg.pipeline).__name__ == "GaussCtrlPipelineConfig"
    assert type(cfg.pipeline.datamanager).__name__ == "GaussCtrlDataManagerConfig"
    assert type(cfg.pipeline.model).__name__ == "GaussCtrlModelConfig"
    dmc = cfg.pipeline.datamanager
    assert (dmc.patch_size, dmc.subset_num, dmc.sampled_views_every_subset, dmc.load_all) == (32, 4, 10, False)
    assert dmc.dataparser.load_3D_points is True

    # ---- the default checkpoint id cannot be resolved offline: loud failure, no silent random weights
    trainer = cfg.setup(local_rank=0, world_size=1)
    try:
        trainer.setup()
        raise SystemExit("expected FileNotFoundError for an unresolvable diffusion_ckpt")
    except FileNotFoundError as exc:
        assert "synthetic" in str(exc)

    # ---- explicit synthetic weights (tiny stand-ins: no GPU here), hot-path calls recorded instead of executed
    tiny = {f"down_blocks.{i}.resnets.0.conv1.weight": torch.zeros(8 * (i + 1), 4, 3, 3) for i in range(4)}
    gp.synthetic_weights = lambda seed: (tiny, tiny, None)
    calls = []
    gp.GaussCtrlPipeline.render_reverse = lambda self: calls.append("render_reverse")
    gp.GaussCtrlPipeline.edit_images = lambda self: calls.append("edit_images")
    cfg.pipeline.diffusion_ckpt = "synthetic"
    cfg.pipeline.render_rate = 7
    trainer = cfg.setup(local_rank=0, world_size=1)
    trainer.setup()
    assert calls == ["render_reverse", "edit_images"]                     # gc_trainer.py:75-78
    pipe = trainer.pipeline
    dm = pipe.datamanager
    assert type(dm).__name__ == "GaussCtrlDataManager" and type(pipe.model).__name__ == "GaussCtrlModel"
    # 96 source images > 4 x 10 -> 40 sampled views, 10 sorted picks from each quarter (gc_datamanager.py:95-111)
    assert len(dm.cameras) == 40 and len(dm.train_data) == 40 and dm.train_unseen_cameras == list(range(40))
    assert [d["image_idx"] for d in dm.train_data] == list(range(40))
    for q in range(4):
        part = dm.sample_idx[10 * q:10 * q + 10]
        assert part == sorted(part) and all(24 * q <= v < 24 * (q + 1) for v in part) and len(set(part)) == 10
    assert pipe.ref_indices == [4, 11, 29, 31] and pipe.num_ref_views == 4   # seed 13789, V=40, R=4
    # VanillaPipeline's contract reached the model: seed points from the dataparser, scene box, num_train_data
    assert pipe.model.seed_points is not None and pipe.model.means.shape == (123, 3) and pipe.model.num_train_data == 96
    cam = pipe._camera_at(3)
    assert cam.shape == (1,)
    trainer.train()
    assert trainer.steps_run == list(range(7))                             # render_rate iterations (gc_trainer.py:186-187)
    camera, data = dm.next_train(0)
    assert camera.shape == (1,) and camera.metadata["cam_idx"] == data["image_idx"] and len(dm.train_unseen_cameras) == 39

    # ---- get_outputs side effects (gc_model.py:84-85, 96-97, 140, 159-160, 170) with the rasteriser stubbed out
    import gaussctrl_b200.gc_model as gm
    from nerfstudio.model_components import renderers
    seen = {}

    def fake_render(params, c2w, fx, fy, cx, cy, H, W, n, background, training=False, state=None):
        seen.update(fx=fx, cx=cx, H=H, W=W, n=n, background=background.clone(), training=training)
        xys = (params["means"][:, :2] * 1.0)          # non-leaf, requires grad in training
        state["xys"], state["radii"] = xys, torch.ones(params["means"].shape[0], dtype=torch.int32)
        return {"rgb": xys.sum() + torch.zeros(H, W, 3), "depth": None, "accumulation": torch.zeros(H, W, 1)}

    gm.render_gaussians = fake_render
    model = pipe.model
    model.step = 0                     # resolution schedule: downscale factor 4 while training at step 0
    model.train()
    cam1 = dm.cameras[0]
    fx0, w0 = float(cam1.fx.item()), int(cam1.width.item())
    out = model.get_outputs(cam1)
    assert seen["training"] and seen["W"] == w0 // 4 and abs(seen["fx"] - fx0 / 4) < 1e-6 and model.last_size == (w0 // 4,) * 2
    assert float(cam1.fx.item()) == fx0 and int(cam1.width.item()) == w0          # camera restored (:170)
    assert seen["n"] == 0                                                          # min(step // interval, sh_degree)
    out["rgb"].sum().backward()
    model.after_train(0)                                                           # asserts xys.grad is not None
    assert model.radii is not None
    renderers.BACKGROUND_COLOR_OVERRIDE = torch.tensor([1.0, 0.0, 1.0])
    model.step = 30000
    outs = model.get_outputs_for_camera(cam1)
    assert torch.equal(seen["background"], torch.tensor([1.0, 0.0, 1.0])) and not seen["training"] and seen["W"] == w0
    assert seen["n"] == 3 and model.training is True and set(outs) == {"rgb", "depth", "accumulation"}
    renderers.BACKGROUND_COLOR_OVERRIDE = None
    model.get_outputs_for_camera(cam1)
    assert torch.equal(seen["background"], model.background_color)
    print("PLUGIN-SEAM-OK")


if __name__ == "__main__":
    main()
