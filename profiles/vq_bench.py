"""The fused VQ kernel alone (batch 64 x 512 tokens, L2 flushed, median of 10) + sha1 of all 32 768 ids: python profiles/vq_bench.py
Used for same-box A/B runs of kernels_simt.cu variants (SELFTOK_B200_LIB=<variant .so>): the ids must be bit-identical."""
import hashlib
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi, config as C, synth  # noqa: E402

dev = torch.device("cuda:0")
d = C.FULL
eng = capi.Engine(d, synth.synth_state_dict(d, device=dev), device=dev, precision="fp16")
x0 = synth.synth_tensor("mb.x0", (64, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
tok, outs_q, feats = eng.encode(x0, return_aux=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(13):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids = eng.vq_argmax(feats)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ids = ids[0] if isinstance(ids, (tuple, list)) else ids
print("vq_ms", sorted(ts[3:])[5], "ids_sha1", hashlib.sha1(tok.cpu().numpy().tobytes()).hexdigest()[:16],
      "vq_ids_sha1", hashlib.sha1(ids.cpu().numpy().tobytes()).hexdigest()[:16],
      "outs_q_sha1", hashlib.sha1(outs_q.cpu().numpy().tobytes()).hexdigest()[:16])
