"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference (never copied).

Used by oracle/gen_golden.py (fixture generation) and by the not-gpu tests that pin oracle/selftok_oracle.py
against the real modules when /root/reference is present.  Nothing in the product package, in the `-m gpu`
tests, in smoke() or in bench.py's GPU arm imports this file; /root/reference does not exist on the GPU box.

What has to be shimmed to import the reference in this image:
  * packages absent here: timm (Mlp/Attention/PatchEmbed), easydict, deepspeed, diffusers  -> oracle/ref_shims/
  * hard-coded ``.cuda()`` on the would-be CPU path (sd3/rectified_flow.py:67, sd3/mmdit.py:1042,
    infer/SelftokPipeline.py:252) -> ``Tensor.cuda`` no-op when no GPU is visible.
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("SELFTOK_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mimogpt"))


_imported = False


def import_reference():
    """Put the reference + shims on sys.path (shims only for packages that are really missing)."""
    global _imported
    if _imported:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util
    need_shims = [m for m in ("timm", "easydict", "deepspeed", "diffusers") if importlib.util.find_spec(m) is None]
    if need_shims:
        sys.path.append(_SHIMS)      # appended: a real package always wins
    sys.path.insert(0, REFERENCE_ROOT)
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self            # noqa: E731
        torch.nn.Module.cuda = lambda self, *a, **k: self         # noqa: E731
    _imported = True


@contextlib.contextmanager
def skip_init():
    """Constructing 2.1 B parameters with the reference's default inits takes minutes; every tensor we care
    about is overwritten by load_state_dict afterwards, so the random inits are patched to no-ops."""
    import torch.nn.init as init
    names = ["xavier_uniform_", "xavier_normal_", "kaiming_uniform_", "kaiming_normal_", "normal_", "uniform_",
             "trunc_normal_", "constant_", "zeros_", "ones_"]
    saved = {n: getattr(init, n) for n in names}
    saved_lin = torch.nn.Linear.reset_parameters
    saved_conv = torch.nn.Conv2d.reset_parameters
    saved_tnormal = torch.Tensor.normal_
    try:
        for n in names:
            setattr(init, n, lambda t, *a, **k: t)
        torch.nn.Linear.reset_parameters = lambda self: None
        torch.nn.Conv2d.reset_parameters = lambda self: None
        torch.Tensor.normal_ = lambda self, *a, **k: self
        yield
    finally:
        for n, f in saved.items():
            setattr(init, n, f)
        torch.nn.Linear.reset_parameters = saved_lin
        torch.nn.Conv2d.reset_parameters = saved_conv
        torch.Tensor.normal_ = saved_tnormal


def dims_to_cfg(dims, enc_name="Enc-Qformer-Uni-XL/2", model_name=None):
    """An EasyDict with the same structure as configs/res256/256-eval.yml for `dims` (used for geometries the
    shipped YAMLs cannot express; the FULL geometry goes through the real YAML instead)."""
    from easydict import EasyDict
    model_name = model_name or ("MMDiT_XL_Renderer" if dims.renderer else "MMDiT_XL")
    return EasyDict(dict(
        common=dict(is_eval=True),
        tokenizer=dict(params=dict(
            image_size=dims.latent * 8, k=dims.K,
            stages=",".join(str(s) for s in dims.stages), k_per_stage=",".join(str(s) for s in dims.k_per_stage),
            gradient_checkpointing=False, in_channels=dims.in_channels, encoder_hidden_size=dims.code_dim,
            ema_enc=False, enc_decay=0.99, L2_lr=0.0, two_part_losses=False, diffusion_type="flow",
            noise_schedule_config=dict(schedule="log_norm", parameterization="velocity", force_recon=False, m=0.0, s=1.0),
            enc=enc_name, enable_enc_variable_size=True,
            encoder_config=dict(time_adaln=True, qformer_mode="dual", pre_norm=False, post_norm=True,
                                xavier_init=False, qk_norm=False, attn_mask=False),
            quantizer_config=dict(codebook_size=dims.codebook_size, code_dim=dims.code_dim, w_diversity=1.0,
                                  ema_entropy_ratio=0.8, w_commit=1.0, decay=0.99, dead_code_threshold=0.2,
                                  reset_cluster_size=0.2, smart_react=True, continuous=False, reg=[0.1, 0.3], K=dims.K),
            model=model_name, context_see_xt=True,
            decoder_config=dict(sd3_cond_pooling=None, class_dropout_prob=0.1, train_filter="all", freeze_filter="",
                                init_method=None, time_adaln="pos_emb", **({"repeat": True} if dims.renderer else {})),
        )),
    ))


def register_geometry(dims, tag):
    """Register constructors for a non-shipped geometry in the reference's OWN registries
    (model_zoo.py:239-280), so ImageTokenizer/SelftokPipeline run unmodified on it."""
    import_reference()
    from mimogpt.models.selftok import model_zoo
    from mimogpt.models.selftok.models_ours import QformerEncoder
    from mimogpt.models.selftok.sd3.mmdit import MMDiT, MMDiT_Renderer

    def enc_ctor(**kwargs):
        return QformerEncoder(patch_size=dims.enc_patch, hidden_size=dims.enc_hidden, num_heads=dims.enc_heads,
                              depth=dims.enc_depth, query_dim=dims.enc_qdim, query_heads=dims.enc_qheads,
                              bidirectional=False, **kwargs)

    def dit_ctor(**kwargs):
        cls = MMDiT_Renderer if dims.renderer else MMDiT
        cec = {"target": "torch.nn.Linear", "params": {"in_features": kwargs["encoder_hidden_size"], "out_features": dims.dit_hidden}}
        return cls(pos_embed_scaling_factor=None, pos_embed_offset=None, pos_embed_max_size=dims.dit_pos_max,
                   patch_size=dims.dit_patch, depth=dims.dit_depth, num_patches=dims.dit_pos_max ** 2,
                   adm_in_channels=kwargs["encoder_hidden_size"], context_embedder_config=cec, device="cpu",
                   dtype=torch.float, **kwargs)

    enc_name, dit_name = f"Enc-{tag}", f"MMDiT-{tag}"
    model_zoo.Enc_models[enc_name] = enc_ctor
    model_zoo.DiT_models[dit_name] = dit_ctor
    # image_tokenizer.py:114 only enables the variable-size path when 'Qformer' is in the encoder name
    enc_name_q = f"Enc-Qformer-{tag}"
    model_zoo.Enc_models[enc_name_q] = enc_ctor
    return enc_name_q, dit_name


class _FakeLatentDist:
    def __init__(self, x):
        self._x = x

    def mode(self):
        return self._x


class FakeVAE:
    """Stand-in for diffusers.AutoencoderKL: identity on tensors.  The VAE is outside the measured path
    (SURVEY 8f rank 1); fixtures are taken at the latent boundary, before any VAE call."""

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def encode(self, x, return_dict=False):
        return (_FakeLatentDist(x),)

    def decode(self, x, return_dict=False):
        return (x,)


def build_reference_pipeline(cfg, state_dict, datasize=None):
    """Run the reference's own SelftokPipeline.__init__ (SelftokPipeline.py:154-207) with (a) a fake VAE and
    (b) torch.load returning the in-memory synthetic state dict."""
    import_reference()
    import diffusers
    from mimogpt.infer import SelftokPipeline as SP
    saved_from_pretrained = diffusers.AutoencoderKL.from_pretrained
    saved_sp_vae = SP.AutoencoderKL
    saved_load = torch.load

    class _AK:
        @classmethod
        def from_pretrained(cls, *a, **k):
            return FakeVAE()

    SP.AutoencoderKL = _AK
    torch.load = lambda *a, **k: state_dict
    try:
        with skip_init():
            pipe = SP.SelftokPipeline(cfg=cfg, ckpt_path="<synthetic>", sd3_path="<none>",
                                      datasize=datasize or cfg.tokenizer.params.image_size,
                                      dtype=torch.float32, device="cpu")
    finally:
        SP.AutoencoderKL = saved_sp_vae
        diffusers.AutoencoderKL.from_pretrained = saved_from_pretrained
        torch.load = saved_load
    return pipe
