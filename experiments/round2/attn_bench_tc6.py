"""Isolated timing + correctness of the tcgen05 attention kernels through the kernel-level ABI entry (conversion included in
the call, so the kernel is timed with CUDA events around repeated launches of the SAME converted buffer is not possible here;
instead the engine-level per-class profile is used):   python profiles/attn_bench.py
Prints per-launch attention time at sampler steps 0 / 25 / 49 for SELFTOK_ATTN=tc5 and tc6 engines (batch 64, fp16)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, REPO)
    from selftoktokenizer_b200 import capi, config as C, synth
    dev = torch.device("cuda:0")
    d = C.FULL
    eng = capi.Engine(d, synth.synth_state_dict(d, device=dev), device=dev, precision="fp16")
    B = 64
    x0 = synth.synth_tensor("bench.x0.0", (B, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
    noise = synth.synth_tensor("bench.noise.0", (B, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
    tok = eng.encode(x0)
    eng.set_use_graph(False)
    rec = {}
    for step in (0, 25, 49):
        eng.dit_velocity(tok, noise, step)
        eng.set_profile(True)
        v = eng.dit_velocity(tok, noise, step)
        prof = eng.get_profile()
        eng.set_profile(False)
        rec[f"step{step}_attn_us"] = round(1000.0 * prof["attention"][0] / prof["attention"][1], 1)
        if step == 0:
            rec["v0_checksum"] = float(v.double().abs().sum())
            torch.save(v.cpu(), f"/tmp/v0_{os.environ.get('SELFTOK_ATTN', 'tc6')}.pt")
    eng.decode(tok, noise)
    eng.set_profile(True)
    x = eng.decode(tok, noise)
    prof = eng.get_profile()
    rec["decode50_ms"] = {k: round(v[0], 1) for k, v in prof.items()}
    rec["x_finite"] = bool(torch.isfinite(x).all())
    print(json.dumps(rec))
else:
    for gen in ("tc5", "tc6", "tc5", "tc6"):
        env = dict(os.environ, SELFTOK_ATTN=gen)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(gen, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-2000:], flush=True)
    import torch
    a, b = torch.load("/tmp/v0_tc5.pt"), torch.load("/tmp/v0_tc6.pt")
    print("v0 max-abs difference tc5 vs tc6:", float((a - b).abs().max()), "of", float(a.abs().max()))
