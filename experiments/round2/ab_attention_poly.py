"""Same-box A/B of the tcgen05 attention's polynomial-exp2 share (selftok_k_set_attn_poly), one engine, one process:

    python profiles/ab_attention.py [precision] > gpurun_out/ab_attention.json

For every setting: per-launch attention time at sampler steps 0 / 25 / 49 (S = 768 / mid / 276; CUDA events around every
launch, graphs off), the class totals of a full 50-step decode, and the deviation of the step-0 velocity from the
all-MUFU (pairs = 0) result.  Batch 64, full geometry, synthetic checkpoint.
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
settings = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4, 5, 6, 8]
eng = capi.Engine(d, synth.synth_state_dict(d, device=dev), device=dev, precision=prec)
B = 64
x0 = synth.synth_tensor("bench.x0.0", (B, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
noise = synth.synth_tensor("bench.noise.0", (B, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
tok = eng.encode(x0)
eng.set_use_graph(False)
out = {"precision": prec, "batch": B, "settings": {}}
v_ref = None
for pairs in settings:
    capi.k_set_attn_poly(pairs)
    rec = {}
    for step in (0, 25, 49):
        eng.dit_velocity(tok, noise, step)
        eng.set_profile(True)
        v = eng.dit_velocity(tok, noise, step)
        prof = eng.get_profile()
        eng.set_profile(False)
        kc = int(eng.tables.k[step]) + 1
        rec[f"step{step}"] = {"S": kc + d.n_img, "attention_us_per_launch": 1000.0 * prof["attention"][0] / prof["attention"][1],
                              "ln_us_per_launch": 1000.0 * prof["ln_modulate"][0] / prof["ln_modulate"][1],
                              "gemm_ms": prof["gemm_tcgen05"][0]}
        if step == 0:
            if v_ref is None:
                v_ref = v.clone()
            rec["v0_max_abs_dev_from_first_setting"] = float((v - v_ref).abs().max())
            rec["v0_abs_max"] = float(v.abs().max())
    eng.decode(tok, noise)
    eng.set_profile(True)
    eng.decode(tok, noise)
    prof = eng.get_profile()
    eng.set_profile(False)
    rec["decode50_class_ms"] = {k: round(v[0], 2) for k, v in prof.items()}
    out["settings"][str(pairs)] = rec
    print(json.dumps({str(pairs): rec}), file=sys.stderr, flush=True)
print(json.dumps(out))
