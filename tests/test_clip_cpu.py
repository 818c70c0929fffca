"""Pins oracle/clip_text.py against the real `transformers.CLIPTextModel` (the module the reference's diffusers
pipeline calls for its prompts, requirements.txt:1) and checks the host-side parameter table of the B200 encoder."""
import pytest
import torch


@pytest.mark.parametrize("layers", [2, 12])
def test_oracle_matches_transformers_clip_text_model(layers):
    transformers = pytest.importorskip("transformers")
    from oracle import clip_text as oc
    cfg = transformers.CLIPTextConfig(vocab_size=oc.VOCAB, hidden_size=oc.HIDDEN, intermediate_size=oc.MLP,
                                      num_hidden_layers=layers, num_attention_heads=oc.HEADS,
                                      max_position_embeddings=oc.MAX_POS, hidden_act="quick_gelu", layer_norm_eps=oc.LN_EPS)
    model = transformers.CLIPTextModel(cfg).eval()
    sd = oc.seeded_state_dict(seed=3, layers=layers)
    assert set(sd) == set(model.state_dict()), set(sd) ^ set(model.state_dict())
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, oc.VOCAB, (3, 77), generator=g)
    ids[:, 0] = 49406
    ids[0, 9:] = 49407   # padded prompt, as CLIPTokenizer(padding="max_length") produces
    with torch.no_grad():
        want = model(input_ids=ids, attention_mask=None)[0]
        got = oc.clip_text_forward(sd, ids, layers=layers)
    assert want.shape == (3, 77, 768)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())
    # causal: a token's embedding does not depend on later tokens
    ids2 = ids.clone()
    ids2[:, 40:] = 1234
    with torch.no_grad():
        got2 = oc.clip_text_forward(sd, ids2, layers=layers)
    assert torch.equal(got2[:, :40], got[:, :40]) and not torch.equal(got2[:, 40:], got[:, 40:])


def test_b200_parameter_table_matches_transformers_names():
    from gaussctrl_b200 import clip_text as ct
    from oracle import clip_text as oc
    sd = oc.seeded_state_dict(seed=0, layers=12)
    assert {k: tuple(v.shape) for k, v in sd.items()} == ct.clip_text_shapes()
    with pytest.raises(ValueError):
        bad = dict(sd)
        bad.pop("text_model.final_layer_norm.bias")
        ct.check_clip_state_dict(bad)


def test_prompt_encoder_factory_without_checkpoint(tmp_path):
    """No tokenizer/ + text_encoder/ under the checkpoint folder -> None (the pipeline then keeps its synthetic
    embeddings); a hub name is never resolved (no network)."""
    from gaussctrl_b200.clip_text import make_prompt_encoder
    assert make_prompt_encoder(str(tmp_path), "cuda") is None
    (tmp_path / "tokenizer").mkdir()
    assert make_prompt_encoder(str(tmp_path), "cuda") is None      # text_encoder/model.safetensors still missing
    assert make_prompt_encoder("CompVis/stable-diffusion-v1-4", "cuda") is None
