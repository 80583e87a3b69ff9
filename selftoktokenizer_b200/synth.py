"""State-dict contract of the path + a seeded synthetic checkpoint.

`state_dict_spec(dims)` lists every tensor the engine reads, under the reference's own checkpoint key
names (SelftokPipeline.py:190-195 loads ``encoder.*`` / ``model.*`` keys into
ImageTokenizer).  No pretrained weights can be fetched offline, so parity and benchmarks run on a
seeded synthetic checkpoint.  The generator is pure integer arithmetic (a splitmix-style hash of the
element index, seeded by a CRC of the key name) followed by three individually-rounded fp32 ops, so the
values are bit-identical on CPU and CUDA and on any host — the GPU box regenerates exactly the tensors
the golden fixtures in tests/golden/ were produced with.

Init scales follow SURVEY 8(c): >=2-D weights ~ U(-a, a) with std = gain / sqrt(fan_in) so activations
stay O(1) through all layers and token ids depend on the image; adaLN linears are NON-zero (the
reference zero-inits them, which would hide modulation bugs); LayerNorm affine ~ 1 + small; the codebook
rows are unit-normalised.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import SelftokDims

Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str, float]]"   # name -> (shape, kind, scale)


def state_dict_spec(d: SelftokDims) -> "OrderedDict[str, Tuple[Tuple[int, ...], str, float]]":
    """name -> (shape, kind, std).  kind in {'w','b','ln_w','ln_b','emb','codebook'}."""
    s: "OrderedDict[str, Tuple[Tuple[int, ...], str, float]]" = OrderedDict()

    def lin(prefix, out_f, in_f, gain=1.0, bias_std=0.02):
        s[prefix + ".weight"] = ((out_f, in_f), "w", gain / math.sqrt(in_f))
        s[prefix + ".bias"] = ((out_f,), "b", bias_std)

    H, Q = d.enc_hidden, d.enc_qdim
    P2 = d.enc_patch * d.enc_patch
    # ---- encoder (models_ours.py:43-98,268-301; modules.py:117-163,277-308; models.py:42-56)
    s["encoder.pos_embed"] = ((1, d.enc_pos_max * d.enc_pos_max, H), "emb", 0.5)
    s["encoder.query_tokens"] = ((1, d.K, Q), "emb", 0.02)   # reference init std (models_ours.py:286)
    s["encoder.x_embedder.proj.weight"] = ((H, d.in_channels, d.enc_patch, d.enc_patch), "w", 1.0 / math.sqrt(d.in_channels * P2))
    s["encoder.x_embedder.proj.bias"] = ((H,), "b", 0.02)
    for i in range(d.enc_depth):
        p = f"encoder.blocks.{i}."
        lin(p + "attn.qkv", 3 * H, H)
        lin(p + "attn.to_query_kv", 2 * Q, H)
        lin(p + "attn.query_linear", 3 * Q, Q)
        lin(p + "attn.proj", H, H)
        lin(p + "attn.query_proj", Q, Q)
        lin(p + "mlp.fc1", 4 * H, H)
        lin(p + "mlp.fc2", H, 4 * H)
        lin(p + "q_mlp.fc1", 4 * Q, Q)
        lin(p + "q_mlp.fc2", Q, 4 * Q)
        lin(p + "adaLN_modulation.1", 6 * Q, Q, gain=0.5)
        lin(p + "t_embedder.mlp.0", Q, 256)
        lin(p + "t_embedder.mlp.2", Q, Q)
    s["encoder.final_layer_norm3.weight"] = ((d.code_dim,), "ln_w", 0.05)
    s["encoder.final_layer_norm3.bias"] = ((d.code_dim,), "ln_b", 0.05)
    lin("encoder.quantizer.project_in", d.code_dim, Q)
    s["encoder.quantizer._codebook.embed"] = ((1, d.codebook_size, d.code_dim), "codebook", 1.0)
    # ---- decoder (sd3/mmdit.py:648-838; renderer :1166-1290)
    D = d.dit_hidden
    s["model.context_pos_embed"] = ((1, d.K, D), "emb", 0.5)
    if d.renderer:
        s["model.positional_embedding"] = ((d.n_img, D), "emb", 0.5)
        s["model.mask_token"] = ((1, 1, D), "emb", 0.5)     # `repeat=True` form
    else:
        s["model.pos_embed"] = ((1, d.dit_pos_max * d.dit_pos_max, D), "emb", 0.5)
        s["model.x_embedder.proj.weight"] = ((D, d.in_channels, d.dit_patch, d.dit_patch), "w", 1.0 / math.sqrt(d.in_channels * d.dit_patch ** 2))
        s["model.x_embedder.proj.bias"] = ((D,), "b", 0.02)
    lin("model.t_embedder.mlp.0", D, 256)
    lin("model.t_embedder.mlp.2", D, D)
    lin("model.context_embedder", D, d.code_dim)
    for j in range(d.dit_depth):
        last = j == d.dit_depth - 1
        for blk in ("context_block", "x_block"):
            p = f"model.joint_blocks.{j}.{blk}."
            pre_only = last and blk == "context_block"
            lin(p + "attn.qkv", 3 * D, D)
            if not pre_only:
                lin(p + "attn.proj", D, D)
                lin(p + "mlp.fc1", 4 * D, D)
                lin(p + "mlp.fc2", D, 4 * D)
            lin(p + "adaLN_modulation.1", (2 if pre_only else 6) * D, D, gain=0.5)
            if blk == "context_block":
                # time_adaln == 'pos_emb' gives every context block its own TimestepEmbedder (mmdit.py:432-436)
                lin(p + "t_embedder.mlp.0", D, 256)
                lin(p + "t_embedder.mlp.2", D, D)
    lin("model.final_layer.linear", d.dit_patch ** 2 * d.in_channels, D)
    lin("model.final_layer.adaLN_modulation.1", 2 * D, D, gain=0.5)
    return s


def _hash_uniform(n: int, seed: int, device) -> torch.Tensor:
    """n fp32 values in [-0.5, 0.5): splitmix64-style integer mixing of (index, seed); exact on CPU and CUDA."""
    out = torch.empty(n, dtype=torch.float32, device=device)
    CH = 1 << 22 if str(device) == "cpu" else 1 << 26     # cache-resident chunks on the host
    add = (seed * 2654435761 + 0x1234567) & 0x3FFFFFFFFFFFFFFF
    base = torch.arange(0, min(CH, n), dtype=torch.int64, device=device)
    tmp = torch.empty_like(base)
    for lo in range(0, n, CH):
        hi = min(n, lo + CH)
        m = hi - lo
        z = base[:m] + lo
        t = tmp[:m]
        z.mul_(-7046029254386353131).add_(add)                               # 0x9E3779B97F4A7C15 as int64
        torch.bitwise_right_shift(z, 30, out=t); t.bitwise_and_(0x3FFFFFFFF); z.bitwise_xor_(t)
        z.mul_(-4658895280553007687)                                         # 0xBF58476D1CE4E5B9
        torch.bitwise_right_shift(z, 27, out=t); t.bitwise_and_(0x1FFFFFFFFF); z.bitwise_xor_(t)
        z.mul_(-7723592293110705685)                                         # 0x94D049BB133111EB
        torch.bitwise_right_shift(z, 31, out=t); t.bitwise_and_(0x1FFFFFFFF); z.bitwise_xor_(t)
        z.bitwise_right_shift_(40).bitwise_and_(0xFFFFFF)                    # 24 random bits, exact in fp32
        u = z.to(torch.float32)
        u.mul_(1.0 / 16777216.0).sub_(0.5)                                   # both exact / individually rounded
        out[lo:hi] = u
    return out


def synth_tensor(name: str, shape, kind: str, std: float, seed: int = 0, device="cpu") -> torch.Tensor:
    n = 1
    for x in shape:
        n *= int(x)
    key = zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1 & 0xFFFFFFFF)
    u = _hash_uniform(n, key, device)
    if kind == "ln_w":
        t = u * float(std * math.sqrt(12.0)) + 1.0
    elif kind == "codebook":
        # unit rows (CosineSimCodebook keeps l2norm'd embeds).  The norm is accumulated column by column in
        # fp64 with elementwise ops only (fixed order; sqrt/div are correctly rounded everywhere), so the
        # codebook is bit-identical on any host/device — no vectorised-reduction order dependence.
        t = u.view(-1, shape[-1]).to(torch.float64)
        ss = torch.zeros(t.shape[0], dtype=torch.float64, device=t.device)
        for j in range(t.shape[1]):
            ss = ss + t[:, j] * t[:, j]
        t = t / torch.sqrt(ss)[:, None]
        return t.to(torch.float32).view(shape).contiguous()
    else:
        t = u * float(std * math.sqrt(12.0))       # uniform with the requested std
    return t.view(shape).contiguous()


def _stress(name: str, t: torch.Tensor, kind: str, seed: int) -> torch.Tensor:
    """Numerically hostile variant of a decoder GEMM weight (fp16-operand stress fixture): heavy-tailed entries (a u + b u^9
    mix of the uniform hash value with the same std: max/std 1.7 -> 4.0) and, in every `attn.qkv` / `mlp.fc1` matrix,
    four output channels scaled by x30 .. x100 -- the outlier channels real SD3-derived checkpoints are known for.
    Pure elementwise fp32 arithmetic on the hashed values: bit-identical on any host and device."""
    if kind != "w" or not name.startswith("model.joint_blocks."):
        return t
    flat = t.reshape(t.shape[0], -1)
    amax = flat.abs().max().clamp_min(1e-30)              # an exact hashed value: no reduction-order dependence
    u = flat / amax                                       # uniform on [-1, 1]
    u3 = u * u * u
    heavy = 0.2 * u + 2.0128 * (u3 * u3 * u3)             # 0.2 u + 0.8 sqrt(19/3) u^9: same std (0.95x), max/std 1.7 -> 4.0
    out = heavy * amax
    if name.endswith("attn.qkv.weight") or name.endswith("mlp.fc1.weight"):
        key = zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1 & 0xFFFFFFFF)
        rows = out.shape[0]
        for i in range(4):
            r = (key * (2 * i + 3) + 977 * i) % rows
            gain = 30.0 + float((key >> (4 * i)) % 71)
            out[r] = out[r] * gain
    return out.reshape(t.shape).contiguous()


def synth_state_dict(d: SelftokDims, seed: int = 0, device="cpu", include_aux: bool = True,
                     stress: bool = False) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for name, (shape, kind, std) in state_dict_spec(d).items():
        sd[name] = synth_tensor(name, shape, kind, std, seed, device)
        if stress:
            sd[name] = _stress(name, sd[name], kind, seed)
    if include_aux:
        # buffers the reference's eval forward consults (vector_quantize_pytorch.py:421-444,555,864-866):
        # `initted`=1 (else eval runs k-means), `continuous`=0.
        sd["encoder.quantizer._codebook.initted"] = torch.ones(1, device=device)
        sd["encoder.quantizer.continuous"] = torch.zeros(1, device=device)
    return sd


# ---------------------------------------------------------------------------------------------- SD3 VAE (16 channels)
def vae_state_dict_spec(ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_channels: int = 16,
                        decoder: bool = True, encoder: bool = True) -> "OrderedDict[str, Tuple[Tuple[int, ...], str, float]]":
    """Tensors of the SD3 VAE under the in-tree `SDVAE` key names (sd3/sd3_impls.py:221-474): GroupNorm(32) + SiLU +
    3x3 convolutions, one single-head attention block in the middle.  ch = 128 is the released architecture."""
    s: "OrderedDict[str, Tuple[Tuple[int, ...], str, float]]" = OrderedDict()

    def conv(prefix, cout, cin, k, gain=1.4):
        s[prefix + ".weight"] = ((cout, cin, k, k), "w", gain / math.sqrt(cin * k * k))
        s[prefix + ".bias"] = ((cout,), "b", 0.02)

    def norm(prefix, c):
        s[prefix + ".weight"] = ((c,), "ln_w", 0.05)
        s[prefix + ".bias"] = ((c,), "ln_b", 0.05)

    def resnet(prefix, cin, cout):
        norm(prefix + ".norm1", cin)
        conv(prefix + ".conv1", cout, cin, 3)
        norm(prefix + ".norm2", cout)
        conv(prefix + ".conv2", cout, cout, 3, gain=0.7)
        if cin != cout:
            conv(prefix + ".nin_shortcut", cout, cin, 1, gain=1.0)

    def attn(prefix, c):
        norm(prefix + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(prefix + "." + n, c, c, 1, gain=1.0)

    nres = len(ch_mult)
    if encoder:
        p = "encoder."
        conv(p + "conv_in", ch, 3, 3, gain=1.0)
        in_mult = (1,) + tuple(ch_mult)
        cin = ch
        for lvl in range(nres):
            cin, cout = ch * in_mult[lvl], ch * ch_mult[lvl]
            for b in range(num_res_blocks):
                resnet(f"{p}down.{lvl}.block.{b}", cin, cout)
                cin = cout
            if lvl != nres - 1:
                conv(f"{p}down.{lvl}.downsample.conv", cin, cin, 3, gain=1.0)
        resnet(p + "mid.block_1", cin, cin)
        attn(p + "mid.attn_1", cin)
        resnet(p + "mid.block_2", cin, cin)
        norm(p + "norm_out", cin)
        conv(p + "conv_out", 2 * z_channels, cin, 3)
    if decoder:
        p = "decoder."
        cin = ch * ch_mult[-1]
        conv(p + "conv_in", cin, z_channels, 3, gain=1.0)
        resnet(p + "mid.block_1", cin, cin)
        attn(p + "mid.attn_1", cin)
        resnet(p + "mid.block_2", cin, cin)
        for lvl in reversed(range(nres)):
            cout = ch * ch_mult[lvl]
            for b in range(num_res_blocks + 1):
                resnet(f"{p}up.{lvl}.block.{b}", cin, cout)
                cin = cout
            if lvl != 0:
                conv(f"{p}up.{lvl}.upsample.conv", cin, cin, 3, gain=1.0)
        norm(p + "norm_out", cin)
        conv(p + "conv_out", 3, cin, 3, gain=0.6)
    return s


def synth_vae_state_dict(ch: int = 128, seed: int = 0, device="cpu", **kw) -> Dict[str, torch.Tensor]:
    """Seeded synthetic VAE checkpoint (same integer-hash generator as the tokenizer's): bit-identical on any host."""
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for name, (shape, kind, std) in vae_state_dict_spec(ch, **kw).items():
        sd[name] = synth_tensor("vae." + name, shape, kind, std, seed, device)
    return sd


def num_params(d: SelftokDims) -> int:
    n = 0
    for shape, _, _ in state_dict_spec(d).values():
        m = 1
        for x in shape:
            m *= x
        n += m
    return n
