"""The nerfstudio method-plugin seam driven end to end against a FAKE `nerfstudio` package that
enforces nerfstudio 1.0.0's contracts (tests/fake_nerfstudio): entry point -> gaussctrl.gc_config.gaussctrl_method ->
Trainer.setup -> VanillaPipeline.__init__ (datamanager / model built from their configs) -> GaussCtrlPipeline, plus the
side effects of GaussCtrlModel.get_outputs that splatfacto's callbacks read.  Runs in a subprocess so this process keeps
the stand-in branch of gaussctrl_b200._compat."""
import os
import subprocess
import sys

from conftest import REPO


def test_plugin_seam_against_fake_nerfstudio():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "tests", "fake_nerfstudio"), REPO, env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "plugin_seam_script.py")], cwd=REPO, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PLUGIN-SEAM-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_shim_modules_reexport_the_reference_names():
    """Module paths and class names of the reference package resolve without nerfstudio installed."""
    import gaussctrl.gc_datamanager as d
    import gaussctrl.gc_model as m
    import gaussctrl.gc_pipeline as p
    import gaussctrl.utils as u
    assert p.GaussCtrlPipeline and p.GaussCtrlPipelineConfig and m.GaussCtrlModel and m.GaussCtrlModelConfig
    assert d.GaussCtrlDataManager and u.CrossViewAttnProcessor and u.compute_attn and u.read_depth2disparity
    c = d.GaussCtrlDataManagerConfig()
    assert (c.patch_size, c.subset_num, c.sampled_views_every_subset, c.load_all) == (32, 4, 10, False)
    assert isinstance(p.GaussCtrlPipelineConfig().datamanager, d.GaussCtrlDataManagerConfig)


def test_unresolvable_checkpoint_raises_and_synthetic_is_explicit():
    """no silent random-weight fallback."""
    import pytest
    from gaussctrl_b200.gc_pipeline import GaussCtrlPipeline, resolve_checkpoint
    with pytest.raises(FileNotFoundError):
        resolve_checkpoint("CompVis/stable-diffusion-v1-4")
    with pytest.raises(FileNotFoundError):
        GaussCtrlPipeline._load_weights("/nonexistent/folder", 0)


def test_view_subset_sampling():
    import random
    from gaussctrl_b200.gc_datamanager import sample_view_subset
    got = sample_view_subset(185, 4, 10, random.Random(3))       # garden: 185 images -> anchors 0,46,92,138 (+185)
    assert len(got) == 40 and len(set(got)) == 40
    bounds = [0, 46, 92, 138, 185]
    for q in range(4):
        part = got[10 * q:10 * q + 10]
        assert part == sorted(part) and all(bounds[q] <= v < bounds[q + 1] for v in part)
