"""Accuracy and time of the six-pass (three-way split bf16) tensor-core GEMM against fp64, next to the fp32 FFMA kernel and the
three-pass split: python profiles/gemm_x6_probe.py
Shapes: the encoder's query-stream linears at batch 64 (M = 32768; K, N = 512/1536, 512/512, 512/2048, 2048/512)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (M, K, N) in [(32768, 512, 1536), (32768, 512, 512), (32768, 512, 2048), (32768, 2048, 512), (4096, 64, 192)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev) * 0.1
    ref = (A.double() @ W.double().t() + b.double())
    scale = float(ref.abs().max())
    row = [f"M={M} K={K} N={N}"]
    for name, fn in (("ffma", lambda: capi.k_linear_f32(A, W, b)), ("x3", lambda: capi.k_linear_tc(A, W, b, 3)),
                     ("x6", lambda: capi.k_linear_tc(A, W, b, 6))):
        out = fn()
        err = float((out.double() - ref).abs().max())
        rel = float(((out.double() - ref).abs() / (ref.abs() + 1e-3 * scale)).max())
        row.append(f"{name}: max-abs {err:.2e} (|y|max {scale:.1f}) rel {rel:.2e} ms {timed(fn):.3f}")
    print(" | ".join(row), flush=True)
print("(the tensor-core times include the fp32 -> plane split of A and W and four cudaMalloc / cudaFree per call)")
