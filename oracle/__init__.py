"""CPU oracle for the GaussCtrl hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch / numpy on the CPU, the algorithm of the reference's hot path
(`gaussctrl/utils.py`, `gaussctrl/gc_model.py`, `gaussctrl/gc_pipeline.py` and the third-party kernels
they call).  It is the *checker* for the CUDA product in `gaussctrl_b200/` and the CPU baseline of
`bench.py`; it is never the thing shipped or measured as the product.

Import policy: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs (`cpu_baseline`,
`--impl reference`, and `extra.reference_gpu_eager` = these same modules run in eager fp16 on the GPU as the
"reference GPU pipeline" denominator) may import anything from here - always as the checker or the baseline, never
inside a timed product region.  Nothing under `gaussctrl_b200/` imports `oracle`.

Parity pinning status:
  * `crossview_attn`  – PINNED: checked against outputs of the reference's own `gaussctrl/utils.py`
                        executed in this container (fixtures in tests/golden/, generator committed).
  * `glue`            – PINNED for depth2disparity / depth2disparity_torch / ref-index selection
                        (reference functions executed via AST extraction; fixtures in tests/golden/).
  * `gsplat_ref`      – PARITY UNPINNED: gsplat 0.1.3 is an un-vendored dependency (README.md:59-60) that is
                        not installed here; its published algorithm is restated from the paper/kernels'
                        documented semantics.
  * `sd15`            – PARITY UNPINNED: diffusers 0.26.0 (requirements.txt:2) is not installed and no
                        checkpoints exist; architecture restated with diffusers-compatible state_dict keys.
"""
