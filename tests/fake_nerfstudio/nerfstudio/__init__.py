"""FAKE `nerfstudio` package for tests/test_plugin_seam_cpu.py (nerfstudio 1.0.0 is not installable in this image).

It carries no rendering or training logic; it ENFORCES the contracts the GaussCtrl plugin relies on, written from
nerfstudio 1.0.0's public behaviour: `InstantiateConfig.setup`, `VanillaPipeline.__init__` building datamanager and
model from their configs (seed points, scene box, num_train_data), `Cameras` indexing / `rescale_output_resolution`,
`FullImageDatamanager` caching, `Trainer.setup/train`, `MethodSpecification`."""
