"""CPU-side checks of the drop-in CrossViewAttnProcessor seam and of the kernel-native golden fixtures
(tests/golden/crossview_native_reference.npz = outputs of the REFERENCE's gaussctrl/utils.py, see make_golden.py)."""
import inspect
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from native_cases import NATIVE_CASES, fingerprint, native_case_inputs, subsample  # noqa: E402

from oracle import crossview_attn as cva  # noqa: E402


@pytest.mark.parametrize("name", sorted(NATIVE_CASES))
def test_oracle_matches_reference_at_native_shapes(name):
    """The oracle's literal restatement reproduces the reference's output bit for bit at the kernels' native shapes
    (and the regenerated inputs are the ones the golden was made from)."""
    z = np.load(os.path.join(GOLDEN, "crossview_native_reference.npz"))
    sd, hs, ehs, (heads, dh, n, f, coeff, cross, ntext) = native_case_inputs(name)
    assert fingerprint(sd, hs, ehs) == bytes(z[f"{name}.fingerprint"]).decode(), "seeded inputs drifted: regenerate goldens"
    attn = cva.AttentionStub(heads * dh, heads, dh, cross_attention_dim=cross)
    attn.load_state_dict(sd)
    with torch.no_grad():
        got = cva.crossview_attention_literal(attn, hs, ehs, coeff)
    sub, mean = subsample(got)
    assert np.array_equal(sub, z[f"{name}.out_sub"])
    assert np.array_equal(mean, z[f"{name}.out_mean"])


def test_processor_signature_matches_reference():
    """gaussctrl/utils.py:39-51: __init__(self_attn_coeff, unet_chunk_size=2); __call__(attn, hidden_states,
    encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0)."""
    from gaussctrl_b200.utils import CrossViewAttnProcessor
    init = inspect.signature(CrossViewAttnProcessor.__init__)
    assert list(init.parameters) == ["self", "self_attn_coeff", "unet_chunk_size"]
    assert init.parameters["unet_chunk_size"].default == 2
    call = inspect.signature(CrossViewAttnProcessor.__call__)
    assert list(call.parameters) == ["self", "attn", "hidden_states", "encoder_hidden_states", "attention_mask", "temb",
                                     "scale"]
    assert [call.parameters[k].default for k in ("encoder_hidden_states", "attention_mask", "temb", "scale")] == \
        [None, None, None, 1.0]
    p = CrossViewAttnProcessor(self_attn_coeff=0.6)
    assert p.self_attn_coeff == 0.6 and p.unet_chunk_size == 2


def test_processor_has_no_cpu_path():
    from gaussctrl_b200._lib import GcbError
    from gaussctrl_b200.utils import CrossViewAttnProcessor
    attn = cva.AttentionStub(80, 2, 40)
    with pytest.raises((GcbError, RuntimeError, AssertionError)):
        CrossViewAttnProcessor(0.6)(attn, torch.randn(10, 16, 80))


def test_processor_raises_indexerror_below_four_frames():
    """The reference's `key[:, [3]*video_length]` raises IndexError when a CFG half has fewer than 4 frames
    (SURVEY §8a gotcha 1); the drop-in keeps that (checked before any kernel is launched)."""
    from gaussctrl_b200.utils import CrossViewAttnProcessor
    attn = cva.AttentionStub(80, 2, 40)
    with pytest.raises(IndexError):
        CrossViewAttnProcessor(0.6)(attn, torch.randn(4, 16, 80))  # F = 2 < 4
