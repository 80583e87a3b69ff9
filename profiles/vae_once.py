"""One batch through both halves of the device VAE (for ncu launch lists): python profiles/vae_once.py [B]
Run under `ncu --profile-from-start off ...`: only the decode + encode after the warm-up are inside the profiled range."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import synth  # noqa: E402
from selftoktokenizer_b200.capi import VaeDecoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
vae = VaeDecoder(synth.synth_vae_state_dict(ch=128, device=dev), device=dev)
z = synth.synth_tensor("bench.noise.0", (B, 16, 32, 32), "emb", 0.5, device=dev)
img = synth.synth_tensor("bench.images", (B, 3, 256, 256), "emb", 0.5, device=dev)
vae.decode(z, norm_ip=True)
vae.encode(img)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
vae.decode(z, norm_ip=True)
e[1].record()
vae.encode(img)
e[2].record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"vae_decode_ms {e[0].elapsed_time(e[1]):.2f} vae_encode_ms {e[1].elapsed_time(e[2]):.2f} (B = {B}; times under ncu are not bench values)")
