"""GPU parity of the VAE encode-mean / decode path (gc_pipeline.py:239-246 and the decode inside pipe(), :209-219)
against oracle/sd15.AutoencoderKL with the same seeded weights (oracle in fp32 on the GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vaes():
    from oracle import sd15
    from gaussctrl_b200.vae import VaeB200
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    ref = sd15.AutoencoderKL().eval()
    for p in ref.parameters():
        p.requires_grad_(False)
    mine = VaeB200(ref.state_dict(), "cuda")
    return ref.cuda(), mine


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_decode_matches_oracle(vaes):
    from gaussctrl_b200 import ops
    ref, mine = vaes
    g = torch.Generator().manual_seed(1)
    z = torch.randn((2, 4, 32, 32), generator=g).half()
    with torch.no_grad():
        want = ref.decode(z.float().cuda())
    got = ops.nhwc_to_nchw(mine.decode(ops.nchw_to_nhwc(z.cuda())))
    rel = _rel(got, want)
    assert rel < 1e-2, rel  # fp16 activations through ~30 conv/GroupNorm layers vs fp32 oracle
    # decode_latents = decode(z / 0.18215) -> (x/2+0.5).clamp(0,1) -> [V,H,W,3] fp32, with and without mask composite
    lat = (z.float() * 0.18215).half()
    imgs = mine.decode_latents(lat.cuda())
    want_img = (ref.decode(lat.float().cuda() / 0.18215) / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)
    assert (imgs - want_img).abs().max().item() < 2e-2
    mask = (torch.rand((2, 256, 256), generator=g) > 0.5).float()
    uned = torch.rand((2, 256, 256, 3), generator=g).half()
    comp = mine.decode_latents(lat.cuda(), mask.cuda(), uned.cuda())
    want_c = want_img.cpu() * mask[..., None] + uned.float() * (1 - mask[..., None])
    assert (comp.cpu() - want_c).abs().max().item() < 2e-2


def test_encode_mean_matches_oracle(vaes):
    ref, mine = vaes
    g = torch.Generator().manual_seed(2)
    img = torch.rand((1, 256, 256, 3), generator=g).half()
    with torch.no_grad():
        want = ref.encode_mean((img.float().cuda() * 2 - 1).permute(0, 3, 1, 2)) * 0.18215
    got = mine.encode_mean(img.cuda())
    rel = _rel(got, want)
    assert rel < 1e-2, rel
