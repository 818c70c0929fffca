"""Pin the oracle against outputs of the reference's own code (fixtures: tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import crossview_attn as cva
from oracle import pipeline as opipe
from conftest import GOLDEN

CASES = ["unet_R4c1", "unet_R4c3", "cnet_R4c3", "unet_d80", "text_cross"]


def _load_case(z, name):
    heads, dh, n, f, cross, ntext = [int(v) for v in z[f"{name}.meta"]]
    attn = cva.AttentionStub(heads * dh, heads, dh, cross_attention_dim=cross or None)
    sd = {k[len(name) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + ".w.")}
    attn.load_state_dict(sd)
    hs = torch.from_numpy(z[f"{name}.hidden"])
    ehs = torch.from_numpy(z[f"{name}.ehs"]) if cross else None
    return attn, hs, ehs, float(z[f"{name}.coeff"][0]), f, torch.from_numpy(z[f"{name}.out"])


@pytest.mark.parametrize("name", CASES)
def test_literal_oracle_matches_reference_output(name):
    z = np.load(os.path.join(GOLDEN, "crossview_reference.npz"))
    attn, hs, ehs, coeff, f, want = _load_case(z, name)
    with torch.no_grad():
        got = cva.crossview_attention_literal(attn, hs, ehs, coeff)
    # same ops in the same order on the same machine: bit-exact
    assert torch.equal(got, want)


@pytest.mark.parametrize("name", CASES[:4])
def test_fused_formulation_matches_reference_output(name):
    z = np.load(os.path.join(GOLDEN, "crossview_reference.npz"))
    attn, hs, ehs, coeff, f, want = _load_case(z, name)
    with torch.no_grad():
        q, k, v = attn.to_q(hs), attn.to_k(hs), attn.to_v(hs)
        ks, vs, ws = cva.crossview_sources(k, v, f, (0, 1, 2, 3), coeff)
        o = cva.multi_source_attention(q, ks, vs, ws, attn.heads)
        got = attn.to_out[0](o)
    assert (got - want).abs().max().item() < 2e-6  # fp32 re-association only


def test_ref_rows_independent_of_chunk_rows():
    """reference-view rows never depend on chunk-view rows."""
    z = np.load(os.path.join(GOLDEN, "crossview_reference.npz"))
    attn, hs, _, coeff, f, want = _load_case(z, "unet_R4c3")
    hs2 = hs.clone()
    g = torch.Generator().manual_seed(5)
    for half in range(2):
        hs2[half * f + 4:(half + 1) * f] = torch.randn(f - 4, *hs.shape[1:], generator=g)
    with torch.no_grad():
        got = cva.crossview_attention_literal(attn, hs2, None, coeff)
    for half in range(2):
        assert torch.equal(got[half * f:half * f + 4], want[half * f:half * f + 4])


def test_glue_matches_reference():
    z = np.load(os.path.join(GOLDEN, "glue_reference.npz"))
    depth = z["depth"]
    assert np.array_equal(opipe.depth2disparity(depth), z["disparity_np"])
    dt = opipe.depth2disparity_torch(torch.from_numpy(depth).to(torch.float16))
    assert np.array_equal(dt.float().numpy(), z["disparity_torch_f16"])
    for key in z.files:
        if key.startswith("ref_indices."):
            _, v, r = key.split(".")
            assert opipe.select_ref_indices(int(v[1:]), int(r[1:])) == z[key].tolist()
    # the documented out-of-range pick for V=1,R=1 (SURVEY §8a gotcha 3) and the clamp the product applies
    assert opipe.select_ref_indices(1, 1) == [1]
    assert opipe.select_ref_indices(1, 1, clamp=True) == [0]
