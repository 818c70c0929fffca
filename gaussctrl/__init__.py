"""`gaussctrl` - the reference's package name, so that `ns-train gaussctrl` (entry point
`gaussctrl.gc_config:gaussctrl_method`, pyproject.toml) resolves to the B200-native implementation of the editing hot
path.  Each module re-exports the class the reference defines under the same module path:

    gaussctrl.gc_pipeline     GaussCtrlPipeline, GaussCtrlPipelineConfig         -> gaussctrl_b200.gc_pipeline
    gaussctrl.gc_model        GaussCtrlModel, GaussCtrlModelConfig               -> gaussctrl_b200.gc_model
    gaussctrl.gc_datamanager  GaussCtrlDataManager, GaussCtrlDataManagerConfig   -> gaussctrl_b200.gc_datamanager
    gaussctrl.utils           CrossViewAttnProcessor, compute_attn, read_depth2disparity -> gaussctrl_b200.utils
    gaussctrl.gc_trainer      GaussCtrlTrainer, GaussCtrlTrainerConfig   (thin subclass of nerfstudio's Trainer)
    gaussctrl.gc_config       gaussctrl_method                           (needs nerfstudio)

Out of the hot path's scope and therefore NOT here: the reference's copy of nerfstudio's dataparser / dataset with
stage-A preload paths (gaussctrl_b200.store reads and writes those folders), the viewer, `ns-gaussctrl-render`, LangSAM."""
