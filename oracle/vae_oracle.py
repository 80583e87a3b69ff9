"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32) of the SD3 16-channel VAE the reference calls on both
sides of the token path (SelftokPipeline.py:162,215,288,316 use diffusers.AutoencoderKL; the same architecture is
vendored in-tree as `SDVAE`, mimogpt/models/selftok/sd3/sd3_impls.py:215-474, which is what BASELINE.md prescribes as
the stand-in for pixel-space parity when no VAE weights exist).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

Parity status: PINNED — tests/test_oracle_pinned.py checks `decode` / `encode_mean` against tests/golden/vae_tiny.npz,
which oracle/gen_golden.py produced by running the unmodified reference `SDVAE` (fp32, CPU) on the seeded synthetic VAE
checkpoint of selftoktokenizer_b200/synth.py, and against the live module when /root/reference is mounted.

State-dict keys are SDVAE's own (`decoder.up.3.block.0.conv1.weight`, `decoder.mid.attn_1.q.weight`, ...).
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _gn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    # Normalize(): GroupNorm(32 groups, eps 1e-6, affine)  (sd3_impls.py:215-218)
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    # ResnetBlock.forward (sd3_impls.py:245-256): norm1 -> swish -> conv1 -> norm2 -> swish -> conv2 (+ 1x1 shortcut)
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h)))
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    # AttnBlock.forward (sd3_impls.py:276-287): single-head attention over the h*w positions, channel dim as head dim
    h = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, p + "." + n, h, padding=0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2).unsqueeze(1) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.squeeze(1).transpose(1, 2).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", o, padding=0)


def decode(sd: SD, z: torch.Tensor, ch_mult: Sequence[int] = (1, 2, 4, 4), num_res_blocks: int = 2) -> torch.Tensor:
    """VAEDecoder.forward (sd3_impls.py:424-444) == SDVAE.decode (:453-455) without autocast (fp32)."""
    p = "decoder."
    h = _conv(sd, p + "conv_in", z.float())
    h = resnet_block(sd, p + "mid.block_1", h)
    h = attn_block(sd, p + "mid.attn_1", h)
    h = resnet_block(sd, p + "mid.block_2", h)
    for lvl in reversed(range(len(ch_mult))):
        for b in range(num_res_blocks + 1):
            h = resnet_block(sd, f"{p}up.{lvl}.block.{b}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")                 # Upsample (sd3_impls.py:309-313)
            h = _conv(sd, f"{p}up.{lvl}.upsample.conv", h)
    return _conv(sd, p + "conv_out", F.silu(_gn(sd, p + "norm_out", h)))


def encode_moments(sd: SD, x: torch.Tensor, ch_mult: Sequence[int] = (1, 2, 4, 4), num_res_blocks: int = 2) -> torch.Tensor:
    """VAEEncoder.forward (sd3_impls.py:365-385): [B,3,H,W] -> [B, 2*z, H/8, W/8] (mean | logvar)."""
    p = "encoder."
    h = _conv(sd, p + "conv_in", x.float())
    for lvl in range(len(ch_mult)):
        for b in range(num_res_blocks):
            h = resnet_block(sd, f"{p}down.{lvl}.block.{b}", h)
        if lvl != len(ch_mult) - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)                   # Downsample (sd3_impls.py:297-301)
            h = _conv(sd, f"{p}down.{lvl}.downsample.conv", h, stride=2, padding=0)
    h = resnet_block(sd, p + "mid.block_1", h)
    h = attn_block(sd, p + "mid.attn_1", h)
    h = resnet_block(sd, p + "mid.block_2", h)
    return _conv(sd, p + "conv_out", F.silu(_gn(sd, p + "norm_out", h)))


def encode_mean(sd: SD, x: torch.Tensor, **kw) -> torch.Tensor:
    """The `.mode()` of the latent distribution the pipeline uses (SelftokPipeline.py:215): the mean half."""
    return encode_moments(sd, x, **kw).chunk(2, dim=1)[0]


# ---- the pixel ends of SelftokPipeline (SelftokPipeline.py:215-218, 284-294) with this VAE as `self.vae`, fp32 ---------
SCALE, SHIFT = 1.5305, 0.0609       # SD3LatentFormat (sd3_impls.py:133-144)


def latents_from_images(sd: SD, images: torch.Tensor) -> torch.Tensor:
    return (encode_mean(sd, images) - SHIFT) * SCALE


def images_from_latents(sd: SD, pred_x0: torch.Tensor) -> torch.Tensor:
    rec = decode(sd, pred_x0.float() / SCALE + SHIFT)
    return (rec.clamp(-1, 1) + 1) / 2                                              # norm_ip(recons, -1, 1)
