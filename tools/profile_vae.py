#!/usr/bin/env python
"""One VAE decode (4 latents) and one VAE encode (4 images) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussctrl_b200.sd15_spec import random_state_dict, vae_shapes
from gaussctrl_b200.vae import VaeB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
vae = VaeB200(random_state_dict(vae_shapes(), 3), "cuda")
lat = torch.randn((B, 4, 64, 64), device="cuda").half()
img = torch.rand((B, 512, 512, 3), device="cuda").half()
for _ in range(2):
    vae.decode_latents(lat, batch=B)
    vae.encode_mean(img)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("profiled")
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
vae.decode_latents(lat, batch=B)
e[1].record()
vae.encode_mean(img)
e[2].record()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(f"decode {e[0].elapsed_time(e[1]) / B:.3f} ms/view, encode {e[1].elapsed_time(e[2]) / B:.3f} ms/view (batch {B})")
