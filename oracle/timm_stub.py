"""A stand-in ``timm`` module (TEST INFRASTRUCTURE -- see oracle/__init__.py) that lets the UNMODIFIED reference
wrappers ``3_Models/backbones/early_fusion_vit.py`` / ``late_fusion_vit.py`` run in this container, where timm itself is
absent (it is an un-vendored, un-pinned dependency of the reference).

``create_model`` returns an object with the attribute surface the wrappers touch -- ``patch_embed.proj`` (a replaceable
``nn.Conv2d``), ``num_features``, ``forward_features``, ``forward`` (``num_classes=0`` -> pooled CLS feature) -- whose
ARITHMETIC is torchvision's ``VisionTransformer`` (an independent implementation of the published ViT: pre-LN blocks,
LayerNorm eps 1e-6, exact GELU, fused in_proj, learned position embedding, class token), NOT oracle/vit.py and NOT the
product.  Weights are exchanged under timm's key names through ``load_timm_state_dict``.
"""
import sys
import types

import torch
import torch.nn as nn
from torchvision.models.vision_transformer import VisionTransformer as _TVViT

VARIANTS = {  # timm registry: (embed_dim, depth, heads)
    "vit_tiny_patch16_224": (192, 12, 3),
    "vit_small_patch16_224": (384, 12, 6),
    "vit_base_patch16_224": (768, 12, 12),
}


class _PatchEmbed(nn.Module):
    """``backbone.patch_embed.proj`` must be an attribute the wrapper can REPLACE (early_fusion_vit.py:114,147)."""

    def __init__(self, proj):
        super().__init__()
        self.proj = proj


class TimmLikeViT(nn.Module):
    def __init__(self, dim, depth, heads, num_classes, img_size):
        super().__init__()
        tv = _TVViT(image_size=img_size, patch_size=16, num_layers=depth, num_heads=heads, hidden_dim=dim,
                    mlp_dim=4 * dim, num_classes=max(num_classes, 1))
        self.patch_embed = _PatchEmbed(tv.conv_proj)
        del tv.conv_proj                                   # single owner: the (replaceable) patch_embed.proj
        self.tv = tv
        self.num_classes = num_classes
        self.num_features = self.embed_dim = dim

    def forward_features(self, x):
        n = x.shape[0]
        t = self.patch_embed.proj(x).flatten(2).transpose(1, 2)            # (B, N, D)
        t = torch.cat([self.tv.class_token.expand(n, -1, -1), t], dim=1)
        return self.tv.encoder(t)                                           # + pos_embedding, blocks, final LN

    def forward(self, x):
        cls = self.forward_features(x)[:, 0]
        return cls if self.num_classes == 0 else self.tv.heads(cls)

    def load_timm_state_dict(self, sd, pre=""):
        """timm key names -> this module (the inverse of tests/test_oracle_vit.py::_tv_to_timm)."""
        tv = {"class_token": sd[pre + "cls_token"], "encoder.pos_embedding": sd[pre + "pos_embed"],
              "encoder.ln.weight": sd[pre + "norm.weight"], "encoder.ln.bias": sd[pre + "norm.bias"]}
        depth = len(self.tv.encoder.layers)
        for i in range(depth):
            s, d = f"{pre}blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
            for a, b in (("norm1", "ln_1"), ("norm2", "ln_2"), ("attn.proj", "self_attention.out_proj"),
                         ("mlp.fc1", "mlp.0"), ("mlp.fc2", "mlp.3")):
                tv[d + b + ".weight"], tv[d + b + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
            tv[d + "self_attention.in_proj_weight"] = sd[s + "attn.qkv.weight"]
            tv[d + "self_attention.in_proj_bias"] = sd[s + "attn.qkv.bias"]
        if self.num_classes > 0:
            tv["heads.head.weight"], tv["heads.head.bias"] = sd[pre + "head.weight"], sd[pre + "head.bias"]
        missing = self.tv.load_state_dict(tv, strict=False)
        assert not missing.unexpected_keys, missing
        assert all(k.startswith("heads.") for k in missing.missing_keys), missing
        with torch.no_grad():
            self.patch_embed.proj.weight.copy_(sd[pre + "patch_embed.proj.weight"])
            self.patch_embed.proj.bias.copy_(sd[pre + "patch_embed.proj.bias"])


def create_model(model_name, pretrained=False, num_classes=1000, img_size=224, **kwargs):
    dim, depth, heads = VARIANTS[model_name]
    return TimmLikeViT(dim, depth, heads, num_classes, img_size).eval()


def install():
    """Registers this module as ``timm`` in sys.modules (only if the real timm is not importable)."""
    try:
        import timm  # noqa: F401
        return sys.modules["timm"]
    except ImportError:
        pass
    mod = types.ModuleType("timm")
    mod.create_model = create_model
    mod.__stub__ = True
    sys.modules["timm"] = mod
    return mod
