"""Per-kernel micro-benchmarks behind the roofline table of DESIGN.md (run on the B200 box; all calls go through the C ABI).

    python profiles/microbench.py > gpurun_out/microbench.json

* VQ (BASELINE config[1]: batch 64, 512 tokens): fused project_in + l2norm + argmax + gather + LN3.  Reported both ways
  (SURVEY 8d): achieved HBM GB/s = 71.6 MB algorithmic bytes / t (what the north-star asks for) and achieved fp32 FFMA
  TFLOP/s = 35.5 GFLOP / t — the kernel is FFMA-bound by construction (arithmetic intensity ~500 FLOP/B).
* tcgen05 GEMM at the four MMDiT shapes (M = 64*768 joint rows), timed in isolation -> compared with the BURST bf16 peak.
* tcgen05 attention at S = 768 and S = 276 (first / last sampler step), B = 64, H = 24.
CUDA events on the launching stream, 3 warm-ups, L2 flushed (256 MiB memset) between timed iterations.
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi, config as C, synth  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


out = {}
# ---- VQ
d = C.FULL
sd = {k: v for k, v in synth.synth_state_dict(d, device=dev).items() if k.startswith("encoder.") or k.startswith("model.")}
eng = capi.Engine(d, sd, device=dev, precision="fp16")
del sd
B = 64
z = torch.randn(B * d.K, d.enc_qdim, device=dev)
ms = timeit(lambda: eng.vq_argmax(z))
alg_bytes = B * d.K * d.enc_qdim * 4 + d.codebook_size * d.code_dim * 4 + d.code_dim * d.enc_qdim * 4 + B * d.K * 8 + B * d.K * d.code_dim * 4
flops = 2.0 * B * d.K * d.codebook_size * d.code_dim + 2.0 * B * d.K * d.enc_qdim * d.code_dim
out["vq"] = {"workload": "batch 64 x 512 tokens, 32768 x 16 codebook, fused project_in", "ms": ms, "algorithmic_MB": alg_bytes / 1e6,
             "achieved_GBps": alg_bytes / ms / 1e6, "hbm_peak_GBps": peaks["hbm_gbs"], "frac_hbm": alg_bytes / ms / 1e6 / peaks["hbm_gbs"],
             "achieved_fp32_TFLOPs": flops / ms / 1e9, "fp32_ffma_peak_TFLOPs_nominal": 148 * 128 * 2 * 1.965e9 / 1e12,
             "note": "FFMA-bound by construction: 35.5 GFLOP per 71.6 MB (SURVEY 8d); the HBM fraction is reported because the north-star asks for it"}
# ---- encode (batch 64)
x0 = synth.synth_tensor("mb.x0", (B, d.in_channels, d.latent, d.latent), "emb", 1.0, device=dev)
out["encode_b64"] = {"ms": timeit(lambda: eng.encode(x0), iters=5), "launches": eng.last_launch_count}
eng.close()
# ---- tcgen05 GEMMs in isolation (A, W converted once outside the timed call is not possible through selftok_k_linear_tc, which
# includes the fp32 -> 16-bit conversion and allocation; so time the engine-level kernels through a profiled decode instead)
eng = capi.Engine(d, synth.synth_state_dict(d, device=dev), device=dev, precision="fp16")
tok = torch.randint(0, d.codebook_size, (B, d.K), device=dev)
noise = torch.randn(B, d.in_channels, d.latent, d.latent, device=dev)
eng.set_use_graph(False)
eng.decode(tok, noise, steps=1)
eng.set_profile(True)
eng.decode(tok, noise, steps=1)
prof = eng.get_profile()
eng.set_profile(False)
D, N, L, Kc = d.dit_hidden, d.n_img, d.dit_depth, d.K
g_flops = sum(B * N * 24 * D * D + B * Kc * (6 * D * D + (0 if j == L - 1 else 18 * D * D)) for j in range(L))
a_flops = L * B * d.dit_heads * 4.0 * (Kc + N) ** 2 * 64
out["decode_step0_b64"] = {k: {"ms": v[0], "launches": v[1]} for k, v in prof.items()}
out["decode_step0_b64"]["gemm_TFLOPs"] = g_flops / prof["gemm_tcgen05"][0] / 1e9
out["decode_step0_b64"]["attention_TFLOPs"] = a_flops / prof["attention"][0] / 1e9
out["decode_step0_b64"]["bf16_peak_sustained"] = peaks["bf16_tflops_sustained"]
eng.close()
print(json.dumps(out, indent=1))
