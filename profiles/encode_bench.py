"""B = 64 encode (patchify + 16 dual blocks + VQ) timed with CUDA events through the C ABI: python profiles/encode_bench.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi, config as C, synth  # noqa: E402

dev = torch.device("cuda:0")
d = C.FULL
sd = {k: v for k, v in synth.synth_state_dict(d, device=dev).items() if k.startswith("encoder.")}
eng = capi.Engine(d, sd, device=dev, precision="fp16", encoder_only=True) if "encoder_only" in capi.Engine.__init__.__code__.co_varnames \
    else capi.Engine(d, synth.synth_state_dict(d, device=dev), device=dev, precision="fp16")
x0 = synth.synth_tensor("mb.x0", (64, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
for _ in range(2):
    tok = eng.encode(x0)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tok = eng.encode(x0)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
tok, _, feats = eng.encode(x0, return_aux=True)
import hashlib
print("encode_b64_ms", sorted(ts)[len(ts) // 2], "tokens_checksum", int(tok.sum().item()),
      "features_sha1", hashlib.sha1(feats.cpu().numpy().tobytes()).hexdigest()[:16])
