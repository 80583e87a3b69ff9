"""Per-class CUDA-event times of one sampler step (default step 0: S = 768, batch 64), graphs off:

    python profiles/step_classes.py [precision] [n_steps] [batch]
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi, config as C, synth  # noqa: E402

dev = torch.device("cuda:0")
d = C.FULL
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = capi.Engine(d, synth.synth_state_dict(d, device=dev), device=dev, precision=prec)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
tok = torch.randint(0, d.codebook_size, (B, d.K), device=dev)
noise = torch.randn(B, d.in_channels, d.latent, d.latent, device=dev)
eng.set_use_graph(False)
eng.decode(tok, noise, steps=steps)
eng.set_profile(True)
out = []
for _ in range(3):
    eng.decode(tok, noise, steps=steps)
    out.append({k: [round(v[0], 3), v[1]] for k, v in eng.get_profile().items()})
print(json.dumps(out[-1]))
print(json.dumps({k: min(o[k][0] for o in out) for k in out[0]}))
if B != 64:                                             # small batches: what the graph replay achieves end to end
    eng.set_profile(False)
    eng.set_use_graph(True)
    eng.decode(tok, noise, steps=steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.decode(tok, noise, steps=steps)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"batch": B, "graph_replay_decode_ms": e0.elapsed_time(e1) / 3, "eager_class_sum_ms": sum(min(o[k][0] for o in out) for k in out[0])}))
