"""Restatement of the three timm 0.9.12 symbols the reference imports
(mimogpt/models/selftok/modules.py:20, models.py:31). Only `Mlp` carries arithmetic on the hot
path (Linear -> act -> Linear; drop=0, norm=Identity); `Attention` is only ever a placeholder
(replaced by DualAttention, modules.py:284) and `PatchEmbed` is imported but the encoder uses the
mmdit.PatchEmbed. TEST INFRASTRUCTURE ONLY - never imported by the product package."""
import torch.nn as nn


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0, use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Attention(nn.Module):
    """Constructed (then discarded) by Encoder.__init__'s placeholder ViTBlocks (models_ours.py:83-85) before
    QformerEncoder replaces `self.blocks` with DualBlocks (models_ours.py:298-301); never executed."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, **kw):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        raise NotImplementedError("timm Attention is never executed on the Selftok hot path")


class PatchEmbed(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()
        raise NotImplementedError("timm PatchEmbed is never instantiated on the Selftok hot path")
