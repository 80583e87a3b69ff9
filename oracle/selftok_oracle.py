"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32) of the reference's encode / decode path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
file; the product package never does (it fails loudly without its CUDA library instead of falling back).

Parity status: PINNED.  tests/test_oracle_pinned.py checks this restatement against tests/golden/*.npz,
which oracle/gen_golden.py produced by running the unmodified reference modules (there are no upstream
golden vectors: SURVEY 4), and — when /root/reference is present — against the live reference modules.

The restatement is a flat functional program over the checkpoint's own key names; every function cites the
reference lines it follows.  The reference is floating point (fp32) end to end, so the oracle is torch fp32
as the tier rules allow for floating-point kernels; torch's F.linear / F.layer_norm / softmax semantics are
the spec the reference itself is written against.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from selftoktokenizer_b200.config import SelftokDims  # noqa: E402
from selftoktokenizer_b200 import schedule as sched  # noqa: E402

SD = Dict[str, torch.Tensor]

# Hook so precision studies can swap the GEMM arithmetic (e.g. emulate split-bf16 tensor-core math).
_linear_impl: Callable = F.linear


def set_linear_impl(fn: Optional[Callable]) -> None:
    global _linear_impl
    _linear_impl = fn or F.linear


def _lin(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return _linear_impl(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(x: torch.Tensor) -> torch.Tensor:
    # nn.LayerNorm(elementwise_affine=False, eps=1e-6): sd3/mmdit.py:386,407,625; modules.py:104-106,286-287
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def _t_embed(sd: SD, prefix: str, freq: torch.Tensor) -> torch.Tensor:
    # TimestepEmbedder.mlp = Linear -> SiLU -> Linear (sd3/mmdit.py:148-152,177-183; models.py:49-53,76-79)
    return _lin(sd, prefix + ".mlp.2", F.silu(_lin(sd, prefix + ".mlp.0", freq)))


def _center_crop_pos(pos: torch.Tensor, max_size: int, h: int, w: int) -> torch.Tensor:
    # cropped_pos_embed (models_ours.py:183-202; sd3/mmdit.py:877-896): centre h x w window of a max x max grid
    top, left = (max_size - h) // 2, (max_size - w) // 2
    g = pos.reshape(1, max_size, max_size, -1)[:, top:top + h, left:left + w, :]
    return g.reshape(1, h * w, -1)


def _patchify(x: torch.Tensor, p: int) -> torch.Tensor:
    # Conv2d(k=p, s=p) as a GEMM over (c, ph, pw)-ordered patch vectors (sd3/mmdit.py:66-75)
    B, Cc, Hh, Ww = x.shape
    x = x.reshape(B, Cc, Hh // p, p, Ww // p, p).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, (Hh // p) * (Ww // p), Cc * p * p)


def _attention(q, k, v, heads, mask=None):
    # other_impls.py:37-45 / modules.py:235-238,263-266: plain SDPA, scale 1/sqrt(head_dim)
    B, Sq, Dm = q.shape
    hd = Dm // heads
    q, k, v = (t.reshape(B, -1, heads, hd).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)
    return o.transpose(1, 2).reshape(B, Sq, Dm)


# ------------------------------------------------------------------------------------------------ encoder

def encoder_features(sd: SD, d: SelftokDims, x0: torch.Tensor, pos_freq: torch.Tensor) -> torch.Tensor:
    """Encoder.forward up to the quantizer input (models_ours.py:204-219) via QformerEncoder.get_encoder_outs
    'dual' mode (models_ours.py:315-343) = enc_depth x DualBlock.forward (modules.py:310-327)."""
    B = x0.shape[0]
    H, Q = d.enc_hidden, d.enc_qdim
    g = d.latent // d.enc_patch
    w = sd["encoder.x_embedder.proj.weight"].reshape(H, -1)
    x = _linear_impl(_patchify(x0.float(), d.enc_patch), w, sd["encoder.x_embedder.proj.bias"])
    x = x + _center_crop_pos(sd["encoder.pos_embed"], d.enc_pos_max, g, g)           # models_ours.py:211-214
    q = sd["encoder.query_tokens"].expand(B, -1, -1)                                  # models_ours.py:316
    for i in range(d.enc_depth):
        p = f"encoder.blocks.{i}."
        # modules.py:311-318: adaLN table from token positions 1000+8k — input independent
        t_emb = _t_embed(sd, p + "t_embedder", pos_freq)
        mod = _lin(sd, p + "adaLN_modulation.1", F.silu(t_emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
        xn = _ln(x)
        qn = _ln(q) * (1 + scale_msa.unsqueeze(0)) + shift_msa.unsqueeze(0)            # modules.py:29-32,321
        # DualAttention uni-directional branch (modules.py:165-174,216-274)
        qkv = _lin(sd, p + "attn.qkv", xn)                                            # [B,N,3H] -> (3, heads, hd)
        xq, xk, xv = qkv.reshape(B, -1, 3, H).unbind(2)
        x_attn = _attention(xq, xk, xv, d.enc_heads)
        kv = _lin(sd, p + "attn.to_query_kv", xn).reshape(B, -1, 2, Q)
        ik, iv = kv.unbind(2)
        qqkv = _lin(sd, p + "attn.query_linear", qn).reshape(B, -1, 3, Q)
        qq, qk, qv = qqkv.unbind(2)
        k_all = torch.cat([ik, qk], dim=1)                                            # modules.py:242-243
        v_all = torch.cat([iv, qv], dim=1)
        q_attn = _attention(qq, k_all, v_all, d.enc_qheads)                           # mask=None (attn_mask: False)
        x_attn = _lin(sd, p + "attn.proj", x_attn)
        q_attn = _lin(sd, p + "attn.query_proj", q_attn)
        x = x + x_attn                                                                # modules.py:322-323
        x = x + _lin(sd, p + "mlp.fc2", _gelu_tanh(_lin(sd, p + "mlp.fc1", _ln(x))))
        q = q + gate_msa.unsqueeze(0) * q_attn                                        # modules.py:325-326
        qm = _ln(q) * (1 + scale_mlp.unsqueeze(0)) + shift_mlp.unsqueeze(0)
        q = q + gate_mlp.unsqueeze(0) * _lin(sd, p + "q_mlp.fc2", _gelu_tanh(_lin(sd, p + "q_mlp.fc1", qm)))
    return q


def vq_argmax(sd: SD, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """VectorQuantize.forward eval branch (vector_quantize_pytorch.py:844-876) -> CosineSimCodebook.forward
    (:525-563,580): project_in, l2norm (:51-52), argmax_c <x_hat, e_c> (first max wins, as torch.argmax),
    gather.  Returns (codes [B,K,dim], ids [B,K] int64)."""
    x = _lin(sd, "encoder.quantizer.project_in", z.float())
    x = F.normalize(x, p=2, dim=-1)
    embed = sd["encoder.quantizer._codebook.embed"][0]
    flat = x.reshape(-1, x.shape[-1])
    ids = torch.empty(flat.shape[0], dtype=torch.long)
    CH = 4096                                            # chunked only to bound memory; per-row result unchanged
    for lo in range(0, flat.shape[0], CH):
        ids[lo:lo + CH] = (flat[lo:lo + CH] @ embed.t()).argmax(dim=-1)
    ids = ids.reshape(x.shape[:-1])
    return embed[ids], ids


def final_norm3(sd: SD, codes: torch.Tensor) -> torch.Tensor:
    # final_layer_norm3: affine LayerNorm(code_dim, eps=1e-6) (models_ours.py:88,241-242)
    return F.layer_norm(codes, (codes.shape[-1],), sd["encoder.final_layer_norm3.weight"],
                        sd["encoder.final_layer_norm3.bias"], eps=1e-6)


def encode(sd: SD, d: SelftokDims, x0: torch.Tensor, tables: Optional[sched.SamplerTables] = None):
    """SelftokPipeline.encoding after the VAE (SelftokPipeline.py:218-225): returns (outs_q, tokens, z)."""
    tables = tables or sched.make_tables(d.K, d.stages, d.k_per_stage)
    z = encoder_features(sd, d, x0, tables.pos_freq)
    codes, ids = vq_argmax(sd, z)
    return final_norm3(sd, codes), ids, z


def lookup(sd: SD, d: SelftokDims, tokens: torch.Tensor) -> torch.Tensor:
    """get_output_from_indices + final_layer_norm3 (SelftokPipeline.py:236-240; vector_quantize_pytorch.py:787-809)."""
    codes = sd["encoder.quantizer._codebook.embed"][0][tokens.long()]
    return final_norm3(sd, codes)


# ------------------------------------------------------------------------------------------------ decoder

def _ctx_adaln(sd: SD, p: str, pos_freq: torch.Tensor) -> torch.Tensor:
    # DismantledBlock.pre_attention, time_adaln == 'pos_emb' (sd3/mmdit.py:446-458): [K, 6D]
    return _lin(sd, p + "adaLN_modulation.1", F.silu(_t_embed(sd, p + "t_embedder", pos_freq)))


def joint_blocks(sd: SD, d: SelftokDims, ctx: torch.Tensor, x: torch.Tensor, c: torch.Tensor,
                 pos_freq: torch.Tensor, n_vis: int, ctx_sees_x: bool, truncate: bool) -> torch.Tensor:
    """forward_core_with_concat (sd3/mmdit.py:918-933): depth x JointBlock -> block_mixing (:508-553), then
    FinalLayer (:641-645).

    n_vis       number of visible context tokens (mask = arange(K) < n_vis; models_ours.py:345-353)
    ctx_sees_x  context rows attend to image keys (MMDiT decode: context_see_xt=True; renderer: False)
    truncate    False: dense K+N joint sequence with the reference's boolean mask (sd3/mmdit.py:1060-1094)
                True : drop context rows >= n_vis (exact: masked as keys everywhere, never read as rows)
    """
    B, Kc, D = ctx.shape
    N = x.shape[1]
    heads = d.dit_heads
    if truncate:
        ctx = ctx[:, :n_vis]
        pos_freq = pos_freq[:n_vis]
        Kc = n_vis
    key_ok = torch.cat([torch.arange(Kc) < n_vis, torch.ones(N, dtype=torch.bool)])          # keys every row may see
    ctx_row = key_ok.clone()
    if not ctx_sees_x:
        ctx_row[Kc:] = False
    mask = torch.cat([ctx_row[None].expand(Kc, -1), key_ok[None].expand(N, -1)], dim=0)[None, None]
    csil = F.silu(c)
    for j in range(d.dit_depth):
        last = j == d.dit_depth - 1
        pc, px = f"model.joint_blocks.{j}.context_block.", f"model.joint_blocks.{j}.x_block."
        # --- pre_attention (sd3/mmdit.py:441-483)
        if not last:
            cm = _ctx_adaln(sd, pc, pos_freq)                                                 # [Kc, 6D]
            c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = cm.chunk(6, dim=1)
            cin = _ln(ctx) * (1 + c_scale_msa.unsqueeze(0)) + c_shift_msa.unsqueeze(0)
        else:                                                                                 # pre_only (:477-483)
            c_shift, c_scale = _lin(sd, pc + "adaLN_modulation.1", csil).chunk(2, dim=1)      # from timestep emb c
            cin = _ln(ctx) * (1 + c_scale.unsqueeze(1)) + c_shift.unsqueeze(1)
        xm = _lin(sd, px + "adaLN_modulation.1", csil)                                        # [B, 6D], 't_emb'
        x_shift_msa, x_scale_msa, x_gate_msa, x_shift_mlp, x_scale_mlp, x_gate_mlp = xm.chunk(6, dim=1)
        xin = _ln(x) * (1 + x_scale_msa.unsqueeze(1)) + x_shift_msa.unsqueeze(1)
        cq, ck, cv = _lin(sd, pc + "attn.qkv", cin).reshape(B, Kc, 3, D).unbind(2)            # split_qkv (:236-238)
        xq, xk, xv = _lin(sd, px + "attn.qkv", xin).reshape(B, N, 3, D).unbind(2)
        # --- joint attention over [context ; x] (:521-531)
        a = _attention(torch.cat([cq, xq], 1), torch.cat([ck, xk], 1), torch.cat([cv, xv], 1), heads, mask)
        c_attn, x_attn = a[:, :Kc], a[:, Kc:]
        # --- post_attention (:485-496)
        if not last:
            ctx = ctx + c_gate_msa.unsqueeze(0) * _lin(sd, pc + "attn.proj", c_attn)
            h = _ln(ctx) * (1 + c_scale_mlp.unsqueeze(0)) + c_shift_mlp.unsqueeze(0)
            ctx = ctx + c_gate_mlp.unsqueeze(0) * _lin(sd, pc + "mlp.fc2", _gelu_tanh(_lin(sd, pc + "mlp.fc1", h)))
        x = x + x_gate_msa.unsqueeze(1) * _lin(sd, px + "attn.proj", x_attn)
        h = _ln(x) * (1 + x_scale_mlp.unsqueeze(1)) + x_shift_mlp.unsqueeze(1)
        x = x + x_gate_mlp.unsqueeze(1) * _lin(sd, px + "mlp.fc2", _gelu_tanh(_lin(sd, px + "mlp.fc1", h)))
    shift, scale = _lin(sd, "model.final_layer.adaLN_modulation.1", csil).chunk(2, dim=1)
    x = _ln(x) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
    return _lin(sd, "model.final_layer.linear", x)


def _unpatchify(x: torch.Tensor, d: SelftokDims) -> torch.Tensor:
    # sd3/mmdit.py:898-916
    B = x.shape[0]
    p, c, g = d.dit_patch, d.in_channels, d.latent // d.dit_patch
    x = x.reshape(B, g, g, p, p, c)
    return torch.einsum("nhwpqc->nchpwq", x).reshape(B, c, g * p, g * p)


def context_embed(sd: SD, outs_q: torch.Tensor) -> torch.Tensor:
    # sd3/mmdit.py:1026: context_embedder(encoder_hidden_states) + context_pos_embed — step invariant
    return _lin(sd, "model.context_embedder", outs_q) + sd["model.context_pos_embed"]


def dit_velocity(sd: SD, d: SelftokDims, x_lat: torch.Tensor, t_freq_row: torch.Tensor, outs_q: torch.Tensor,
                 pos_freq: torch.Tensor, n_vis: int, truncate: bool = False) -> torch.Tensor:
    """MMDiT.forward (sd3/mmdit.py:992-1101) for y=None, cfg_scale 1, eval."""
    B = x_lat.shape[0]
    D = d.dit_hidden
    g = d.latent // d.dit_patch
    w = sd["model.x_embedder.proj.weight"].reshape(D, -1)
    x = _linear_impl(_patchify(x_lat.float(), d.dit_patch), w, sd["model.x_embedder.proj.bias"])
    x = x + _center_crop_pos(sd["model.pos_embed"], d.dit_pos_max, g, g)                       # :1001
    c = _t_embed(sd, "model.t_embedder", t_freq_row.reshape(1, -1)).expand(B, -1)               # :1022 (t equal over batch)
    ctx = context_embed(sd, outs_q)
    out = joint_blocks(sd, d, ctx, x, c, pos_freq, n_vis, ctx_sees_x=True, truncate=truncate)
    return _unpatchify(out, d)


def decode(sd: SD, d: SelftokDims, tokens: torch.Tensor, noise: torch.Tensor, steps: int = 50,
           truncate: bool = False, replay_dead_encoder_call: bool = False,
           tables: Optional[sched.SamplerTables] = None, n_steps_run: Optional[int] = None) -> torch.Tensor:
    """SelftokPipeline.decoding up to pred_x0 (SelftokPipeline.py:227-282) + RectifiedFlow.p_sample_loop
    (rectified_flow.py:165-256) with euler_step (:301-309), cfg_scale == 1.

    replay_dead_encoder_call=True re-runs encoder+VQ on the initial noise every step, as the reference does just
    to obtain `arange(K) <= k` (rectified_flow.py:212-215) — only meaningful when TIMING the reference's CPU cost.
    """
    tb = tables or sched.make_tables(d.K, d.stages, d.k_per_stage, steps)
    outs_q = lookup(sd, d, tokens)
    x = noise.float().clone()
    x_init = x.clone()
    for i in range(steps if n_steps_run is None else n_steps_run):
        if replay_dead_encoder_call:
            encode(sd, d, x_init, tb)
        v = dit_velocity(sd, d, x, tb.t_freq[i], outs_q, tb.pos_freq, int(tb.k[i]) + 1, truncate)
        x = x - tb.dt[i] * v                                                                    # rectified_flow.py:303
    return x


def dit_velocity_uncond(sd: SD, d: SelftokDims, x_lat: torch.Tensor, t_freq_u_row: torch.Tensor) -> torch.Tensor:
    """MMDiT.cfg_inference (sd3/mmdit.py:1117-1163) as sample_one_step calls it (encoder_hidden_states=None, mask = zeros):
    context = zeros and every context key masked for EVERY row -> restated here without truncation (K zero rows kept)."""
    B = x_lat.shape[0]
    D = d.dit_hidden
    g = d.latent // d.dit_patch
    w = sd["model.x_embedder.proj.weight"].reshape(D, -1)
    x = _linear_impl(_patchify(x_lat.float(), d.dit_patch), w, sd["model.x_embedder.proj.bias"])
    x = x + _center_crop_pos(sd["model.pos_embed"], d.dit_pos_max, g, g)
    c = _t_embed(sd, "model.t_embedder", t_freq_u_row.reshape(1, -1)).expand(B, -1)       # integer timestep (mmdit.py:1127-1130)
    ctx = torch.zeros(B, d.K, D)                                                          # mmdit.py:1146 (no pos embed, no embedder)
    pos_freq = sched.make_tables(d.K, d.stages, d.k_per_stage, 1).pos_freq
    # mask = cat(zeros, ones) repeated for every row (mmdit.py:1158-1159): all rows, context rows included, see the image keys only
    out = joint_blocks(sd, d, ctx, x, c, pos_freq, n_vis=0, ctx_sees_x=True, truncate=False)
    return _unpatchify(out, d)


def decode_cfg(sd: SD, d: SelftokDims, tokens: torch.Tensor, noise: torch.Tensor, cfg_scale: float, steps: int = 50) -> torch.Tensor:
    """p_sample_loop with uncond_scale = cfg_scale (rectified_flow.py:165-256) -> sample_one_step's guided branch (:280-289):
    out_uncond = cfg_inference(x, t, None, None, mask = 0); out = model(x, t, None, context, mask = ori_mask) -- WITHOUT
    context_see_xt (default False: context rows only see the visible context keys); v = out_uncond + s (out - out_uncond)."""
    tb = sched.make_tables(d.K, d.stages, d.k_per_stage, steps)
    outs_q = lookup(sd, d, tokens)
    x = noise.float().clone()
    for i in range(steps):
        n_vis = int(tb.k[i]) + 1
        B = x.shape[0]
        D = d.dit_hidden
        g = d.latent // d.dit_patch
        w = sd["model.x_embedder.proj.weight"].reshape(D, -1)
        xe = _linear_impl(_patchify(x, d.dit_patch), w, sd["model.x_embedder.proj.bias"])
        xe = xe + _center_crop_pos(sd["model.pos_embed"], d.dit_pos_max, g, g)
        c = _t_embed(sd, "model.t_embedder", tb.t_freq[i].reshape(1, -1)).expand(B, -1)
        v_c = _unpatchify(joint_blocks(sd, d, context_embed(sd, outs_q), xe, c, tb.pos_freq, n_vis, ctx_sees_x=False, truncate=False), d)
        v_u = dit_velocity_uncond(sd, d, x, tb.t_freq_uncond[i])
        x = x - tb.dt[i] * (v_u + cfg_scale * (v_c - v_u))
    return x


def render(sd: SD, d: SelftokDims, tokens: torch.Tensor, truncate: bool = False) -> torch.Tensor:
    """decoding_with_renderer up to pred_x0 (SelftokPipeline.py:296-310): one MMDiT_Renderer.forward
    (sd3/mmdit.py:1511-1620): x = mask_token + positional_embedding, t = 1000 (no *1000), context rows see context only."""
    B = tokens.shape[0]
    outs_q = lookup(sd, d, tokens)
    x = (sd["model.mask_token"].expand(B, d.n_img, -1) + sd["model.positional_embedding"]).contiguous()
    c = _t_embed(sd, "model.t_embedder", sched.renderer_t_freq()).expand(B, -1)
    ctx = context_embed(sd, outs_q)
    pos_freq = sched.make_tables(d.K, d.stages, d.k_per_stage, 1).pos_freq
    out = joint_blocks(sd, d, ctx, x, c, pos_freq, d.K, ctx_sees_x=False, truncate=truncate)
    return _unpatchify(out, d)


# ------------------------------------------------------------------------------------------------ precision studies

def make_split_bf16_linear(n_terms: int) -> Callable:
    """Emulates the tensor-core GEMM arithmetic of the CUDA path on the CPU: operands rounded to bf16 (hi) and,
    for n_terms == 3, a bf16 residual (lo); products hi*hi (+ hi*lo + lo*hi) accumulated in fp32."""

    def split(t):
        hi = t.to(torch.bfloat16).float()
        lo = (t - hi).to(torch.bfloat16).float()
        return hi, lo

    def lin(x, w, b=None):
        xh, xl = split(x)
        wh, wl = split(w)
        y = F.linear(xh, wh)
        if n_terms == 3:
            y = y + F.linear(xh, wl) + F.linear(xl, wh)
        return y if b is None else y + b

    return lin
