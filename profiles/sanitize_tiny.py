"""compute-sanitizer target: tiny-geometry encode + 6-step decode + renderer-free VAE round trip, eager (graphs off).
    compute-sanitizer --tool memcheck python profiles/sanitize_tiny.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import config as C, synth  # noqa: E402
from selftoktokenizer_b200.capi import Engine, VaeDecoder  # noqa: E402

d = C.TINY
sd = synth.synth_state_dict(d)
for prec in ("fp16", "bf16x3"):
    eng = Engine(d, sd, device="cuda:0", precision=prec, steps=6)
    eng.set_use_graph(False)
    x0 = synth.synth_tensor("smoke.x0", (3, d.in_channels, d.latent, d.latent), "emb", 1.0)
    noise = synth.synth_tensor("smoke.noise", (3, d.in_channels, d.latent, d.latent), "emb", 1.0)
    tok = eng.encode(x0)
    x = eng.decode(tok, noise)
    torch.cuda.synchronize()
    print(prec, "decode finite:", bool(torch.isfinite(x).all()), flush=True)
    eng.close()
vae = VaeDecoder(synth.synth_vae_state_dict(ch=128), device="cuda:0")
z = synth.synth_tensor("smoke.z", (1, 16, 8, 8), "emb", 1.0)
img = synth.synth_tensor("smoke.img", (1, 3, 128, 128), "emb", 0.5)
a, b = vae.decode(z), vae.encode(img)
torch.cuda.synchronize()
print("vae finite:", bool(torch.isfinite(a).all() and torch.isfinite(b).all()), flush=True)
