"""Config surface of the Selftok encode/decode path.

The reference reads only ``cfg.tokenizer.params.*`` and ``cfg.common.is_eval`` at inference
(mimogpt/infer/SelftokPipeline.py:166-168,182-185; mimogpt/models/selftok/image_tokenizer.py:58-159).
``SelftokDims`` is the flat, validated view of those fields that the C-ABI engine consumes
(`include/selftok_b200.h: selftok_config_t`).  The registries map the reference's model names
(mimogpt/models/selftok/model_zoo.py:22-60,177-180,239-280) to the architecture numbers; only the
entries the shipped YAMLs select are present — anything else raises, as the reference's dict lookup would.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, Mapping, Optional, Tuple

import yaml

# model_zoo.py:177-180  Enc_Qformer_Uni_XL_2 -> QformerEncoder(patch 2, hidden 64, 4 heads, depth 16,
#                        query_dim 512, query_heads 8, bidirectional=False)
ENC_MODELS: Dict[str, Dict[str, int]] = {
    "Enc-Qformer-Uni-XL/2": dict(patch=2, hidden=64, heads=4, depth=16, query_dim=512, query_heads=8),
}
# model_zoo.py:22-60: MMDiT_XL / MMDiT_XL_Renderer -> depth 24 (hidden = 64*depth, heads = depth),
# pos_embed_max_size 192, num_patches 36864, patch 2.
DIT_MODELS: Dict[str, Dict[str, Any]] = {
    "MMDiT_XL": dict(depth=24, pos_embed_max_size=192, renderer=False),
    "MMDiT_XL_Renderer": dict(depth=24, pos_embed_max_size=192, renderer=True),
}


class AttrDict(dict):
    """Attribute-access dict with the EasyDict behaviour the reference relies on
    (mimogpt/infer/infer_utils.py:12-19,165-168): nested dicts become AttrDicts, ``cfg.a.b`` works,
    ``hasattr(cfg, 'x')`` is False for missing keys, ``pop`` removes the attribute too."""

    def __init__(self, d: Optional[Mapping] = None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, Mapping) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError:
            raise AttributeError(k)


def parse_args_from_yaml(yml_path: str) -> AttrDict:
    """Same contract as mimogpt/infer/infer_utils.py:165-168 (YAML file -> attribute dict)."""
    with open(yml_path, "r") as fd:
        return AttrDict(yaml.load(fd.read(), Loader=yaml.FullLoader))


def _parse_int_list(s) -> Tuple[int, ...]:
    if isinstance(s, str):
        return tuple(int(x) for x in s.split(","))
    return tuple(int(x) for x in s)


@dataclasses.dataclass(frozen=True)
class SelftokDims:
    # tokens / schedule (256-eval.yml:51-55)
    K: int = 512
    stages: Tuple[int, ...] = (200, 400, 600, 800, 1000)
    k_per_stage: Tuple[int, ...] = (192, 184, 72, 48, 16)
    # latent geometry: image_size // 8 (image_tokenizer.py:111), in_channels 16
    latent: int = 32
    in_channels: int = 16
    # encoder (model_zoo.py:177-180)
    enc_patch: int = 2
    enc_hidden: int = 64
    enc_heads: int = 4
    enc_depth: int = 16
    enc_qdim: int = 512
    enc_qheads: int = 8
    enc_pos_max: int = 64           # pos_embed_max_size = 2*latent (image_tokenizer.py:116-118)
    # quantizer (256-eval.yml:83-85); code_dim == encoder_hidden_size -> project_out Identity
    codebook_size: int = 32768
    code_dim: int = 16
    # decoder (model_zoo.py:22-60; sd3/mmdit.py:708-709)
    dit_depth: int = 24
    dit_patch: int = 2
    dit_pos_max: int = 192
    renderer: bool = False
    context_see_xt: bool = True

    @property
    def dit_hidden(self) -> int:
        return 64 * self.dit_depth

    @property
    def dit_heads(self) -> int:
        return self.dit_depth

    @property
    def n_img(self) -> int:
        return (self.latent // self.dit_patch) ** 2

    @property
    def enc_n_img(self) -> int:
        return (self.latent // self.enc_patch) ** 2

    def validate(self) -> None:
        """Raises ValueError (never a bare assert: the checks must survive `python -O`)."""
        def need(cond, msg):
            if not cond:
                raise ValueError("SelftokDims: " + msg)
        need(self.enc_hidden % self.enc_heads == 0 and self.enc_qdim % self.enc_qheads == 0, "hidden sizes must divide into heads")
        need(self.latent % self.enc_patch == 0 and self.latent % self.dit_patch == 0, "latent side must be a multiple of the patch size")
        need(sum(self.k_per_stage) == self.K, "k_per_stage must sum to K")
        need(len(self.stages) == len(self.k_per_stage), "stages and k_per_stage must have the same length")
        need(self.latent // self.enc_patch <= self.enc_pos_max, "encoder positional grid smaller than the latent grid")
        need(self.latent // self.dit_patch <= self.dit_pos_max, "decoder positional grid smaller than the latent grid")

    @staticmethod
    def from_cfg(cfg: Mapping, datasize: Optional[int] = None) -> "SelftokDims":
        """Flatten ``cfg.tokenizer.params`` exactly as ImageTokenizer.__init__ consumes it
        (image_tokenizer.py:85-147).  Does NOT mutate cfg (the reference does: SelftokPipeline.py:166,
        image_tokenizer.py:88-92 — documented quirk, consciously dropped).  Unsupported settings raise ValueError / KeyError
        (registry lookups), as the reference's constructors would fail."""
        def need(cond, msg):
            if not cond:
                raise ValueError("selftok config: " + msg)
        p = cfg["tokenizer"]["params"]
        enc = ENC_MODELS[p["enc"]]
        dit = DIT_MODELS[p["model"]]
        need(p.get("diffusion_type", "flow") == "flow", "only diffusion_type 'flow' is on the shipped path")
        image_size = int(datasize or p["image_size"])
        need(image_size % 8 == 0, "Image size must be divisible by 8 (for the VAE encoder).")
        latent = image_size // 8
        ec = p.get("encoder_config", {})
        need(ec.get("qformer_mode", "dual") == "dual", "only the 'dual' Q-Former mode is on the shipped path")
        need(ec.get("time_adaln", True) and not ec.get("attn_mask", False) and not ec.get("qk_norm", False),
             "encoder_config must be time_adaln, no attn_mask, no qk_norm")
        need(ec.get("post_norm", True) and not ec.get("pre_norm", False), "encoder_config must be post_norm")
        qc = p["quantizer_config"]
        need(not qc.get("continuous", False), "continuous quantizer is not on the shipped path")
        need(int(qc["code_dim"]) == int(p["encoder_hidden_size"]), "project_out must be Identity (code_dim == encoder_hidden_size)")
        dc = p.get("decoder_config", {})
        need(dc.get("time_adaln", "pos_emb") == "pos_emb", "decoder_config.time_adaln must be 'pos_emb'")
        d = SelftokDims(
            K=int(p["k"]), stages=_parse_int_list(p["stages"]), k_per_stage=_parse_int_list(p["k_per_stage"]),
            latent=latent, in_channels=int(p.get("in_channels", 16)),
            enc_patch=enc["patch"], enc_hidden=enc["hidden"], enc_heads=enc["heads"], enc_depth=enc["depth"],
            enc_qdim=enc["query_dim"], enc_qheads=enc["query_heads"],
            enc_pos_max=2 * (int(p["image_size"]) // 8) if p.get("enable_enc_variable_size", False) else latent // enc["patch"],
            codebook_size=int(qc["codebook_size"]), code_dim=int(qc["code_dim"]),
            dit_depth=dit["depth"], dit_pos_max=dit["pos_embed_max_size"], renderer=dit["renderer"],
            context_see_xt=bool(p.get("context_see_xt", False)),
        )
        d.validate()
        return d


# A reduced geometry that keeps every code path (pos-crop, ragged K, pre_only last layer) but runs the
# reference on CPU in well under a second; used by tests/golden fixtures.
TINY = SelftokDims(
    K=32, stages=(200, 400, 600, 800, 1000), k_per_stage=(12, 8, 6, 4, 2),
    latent=8, in_channels=16,
    enc_patch=2, enc_hidden=64, enc_heads=4, enc_depth=2, enc_qdim=128, enc_qheads=2, enc_pos_max=16,
    codebook_size=1024, code_dim=16,
    dit_depth=3, dit_patch=2, dit_pos_max=12, renderer=False, context_see_xt=True,
)
# A mid-size geometry (multi-tile everywhere: S up to 192 joint rows, 6 heads, 4096 codes) for batched fixtures the CPU
# reference still finishes in seconds: B >= 4 decode and the fp16 outlier-weight stress fixture.
MID = SelftokDims(
    K=128, stages=(200, 400, 600, 800, 1000), k_per_stage=(48, 46, 18, 12, 4),
    latent=16, in_channels=16,
    enc_patch=2, enc_hidden=64, enc_heads=4, enc_depth=4, enc_qdim=256, enc_qheads=4, enc_pos_max=32,
    codebook_size=4096, code_dim=16,
    dit_depth=6, dit_patch=2, dit_pos_max=24, renderer=False, context_see_xt=True,
)
FULL = SelftokDims()
