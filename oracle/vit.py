"""CPU restatement of the gaze encoders (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference builds its ViT with ``timm.create_model(model_name, pretrained, num_classes[, img_size])``
(3_Models/backbones/early_fusion_vit.py:86-91 = ``efv``, late_fusion_vit.py:106-110 = ``lfv``).  timm is a
third-party dependency that is NOT vendored under /root/reference, is not installed in this image and is
not version-pinned by the reference (no requirements / lock file).  PARITY FOR THE ViT ARITHMETIC IS
THEREFORE UNPINNED: what follows restates timm's published ``VisionTransformer`` definition
(``vit_*_patch16_224``): patch-embed Conv2d(k=16,s=16) -> flatten -> [cls] + tokens -> + pos_embed (1,197,D) ->
depth x pre-LN blocks { LN(eps 1e-6) -> fused qkv Linear(D,3D,bias) -> softmax((q*hd^-0.5) k^T) v -> proj ;
LN -> fc1 -> GELU(erf) -> fc2 } -> LN -> CLS token -> head, with timm's state_dict key names.
It is cross-checked against torchvision's VisionTransformer by weight remapping (tests/test_oracle_vit.py).
The wrapper logic (input fusion modes, 6-channel patch-embed surgery, CLS-feature fusion) follows the
reference files directly.
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# timm registry entries the reference configs can name: (embed_dim, depth, heads); mlp_ratio 4, patch 16
VIT_VARIANTS = {
    "vit_tiny_patch16_224": (192, 12, 3),
    "vit_small_patch16_224": (384, 12, 6),
    "vit_base_patch16_224": (768, 12, 12),
    "vit_large_patch16_224": (1024, 24, 16),
}
EARLY_MODES = ["concat", "add", "subtract", "subtract_abs", "multiply"]  # efv:79
LATE_MODES = ["concat", "add", "subtract", "multiply", "full"]           # lfv:98


def vit_features(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, patch: int = 16, taps: dict = None) -> Tensor:
    """timm VisionTransformer.forward_features -> (B, 1+N, D) after the final LayerNorm.  ``taps['last_block']`` receives
    the output of blocks[-1] with its gradient retained (what 6_Utils/attention_utils.py:196-215 hooks)."""
    w = sd[pre + "patch_embed.proj.weight"]
    D = w.shape[0]
    t = F.conv2d(x, w, sd[pre + "patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    B = t.shape[0]
    t = torch.cat([sd[pre + "cls_token"].expand(B, -1, -1), t], dim=1) + sd[pre + "pos_embed"]
    hd = D // heads
    depth = 0
    while f"{pre}blocks.{depth}.norm1.weight" in sd:
        depth += 1
    for i in range(depth):
        p = f"{pre}blocks.{i}."
        h = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        qkv = qkv.reshape(B, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
        a = F.softmax(q @ k.transpose(-2, -1), dim=-1)
        h = (a @ v).transpose(1, 2).reshape(B, -1, D)
        t = t + F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        t = t + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    if taps is not None:
        if t.requires_grad:
            t.retain_grad()
        taps["last_block"] = t
    return F.layer_norm(t, (D,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-6)


def fuse_inputs(a: Tensor, b: Tensor, mode: str) -> Tensor:
    """EarlyFusionViT._fuse_inputs (efv:163-196)."""
    if mode not in EARLY_MODES:
        raise ValueError(f"fusion_mode must be one of {EARLY_MODES}, got '{mode}'")
    if mode == "concat":
        return torch.cat([a, b], dim=1)
    if mode == "add":
        return (a + b) / 2.0
    if mode == "subtract":
        return (a - b) / 2.0
    if mode == "subtract_abs":
        return torch.abs(a - b)
    prod = a * b
    B, C, H, W = prod.shape
    flat = prod.view(B, C, -1)
    flat = (flat - flat.mean(2, keepdim=True)) / (flat.std(2, keepdim=True) + 1e-6)  # unbiased std, efv:191-193
    return flat.view(B, C, H, W)


def early_fusion_forward(sd: Dict[str, Tensor], a: Tensor, b: Tensor, heads: int, mode: str = "concat") -> Tensor:
    """EarlyFusionViT.forward (efv:221-226): logits = head(CLS)."""
    f = vit_features(fuse_inputs(a, b, mode), sd, "backbone.", heads)
    return F.linear(f[:, 0], sd["backbone.head.weight"], sd["backbone.head.bias"])


def early_fusion_features(sd: Dict[str, Tensor], a: Tensor, b: Tensor, heads: int, mode: str = "concat") -> Tensor:
    """EarlyFusionViT.get_features (efv:239-242)."""
    return vit_features(fuse_inputs(a, b, mode), sd, "backbone.", heads)[:, 0]


def fuse_features(c1: Tensor, c2: Tensor, mode: str) -> Tensor:
    """LateFusionViT._fuse_features (lfv:148-178)."""
    if mode not in LATE_MODES:
        raise ValueError(f"fusion_mode must be one of {LATE_MODES}, got '{mode}'")
    if mode == "concat":
        return torch.cat([c1, c2], dim=1)
    if mode == "add":
        return c1 + c2
    if mode == "subtract":
        return c1 - c2
    if mode == "multiply":
        return c1 * c2
    return torch.cat([c1, c2, c1 - c2, c1 * c2], dim=1)


def late_fusion_forward(sd: Dict[str, Tensor], x1: Tensor, x2: Tensor, heads: int, mode: str = "full") -> Tensor:
    """LateFusionViT.forward (lfv:214-228), eval mode."""
    c1 = vit_features(x1, sd, "encoder.", heads)[:, 0]
    c2 = vit_features(x2, sd, "encoder.", heads)[:, 0]
    return F.linear(fuse_features(c1, c2, mode), sd["classifier.weight"], sd["classifier.bias"])


def widen_patch_embed(w3: Tensor, strategy: str = "duplicate") -> Tensor:
    """efv:133-142: 3-channel patch-embed weight -> 6-channel ('duplicate' or 'average')."""
    w6 = torch.empty(w3.shape[0], 6, *w3.shape[2:], dtype=w3.dtype)
    w6[:, 0:3] = w3
    w6[:, 3:6] = w3 if strategy == "duplicate" else w3.mean(dim=1, keepdim=True).expand_as(w3)
    return w6


def init_vit_state_dict(model_name: str, in_chans: int, num_classes: int, pre: str, seed: int = 0,
                        img: int = 224, patch: int = 16) -> Dict[str, Tensor]:
    """Random ViT parameters under timm key names (shapes per variant); ``num_classes=0`` omits the head."""
    D, depth, _ = VIT_VARIANTS[model_name]
    g = torch.Generator().manual_seed(seed)
    n = (img // patch) ** 2
    sd: Dict[str, Tensor] = {}

    def lin(name, o, i, std=0.02):
        sd[name + ".weight"] = std * torch.randn(o, i, generator=g)
        sd[name + ".bias"] = 0.02 * torch.randn(o, generator=g)

    def ln(name):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(D, generator=g)

    sd[pre + "cls_token"] = 0.02 * torch.randn(1, 1, D, generator=g)
    sd[pre + "pos_embed"] = 0.02 * torch.randn(1, n + 1, D, generator=g)
    fan = in_chans * patch * patch
    sd[pre + "patch_embed.proj.weight"] = torch.randn(D, in_chans, patch, patch, generator=g) / math.sqrt(fan)
    sd[pre + "patch_embed.proj.bias"] = 0.02 * torch.randn(D, generator=g)
    for i in range(depth):
        p = f"{pre}blocks.{i}."
        ln(p + "norm1")
        lin(p + "attn.qkv", 3 * D, D)
        lin(p + "attn.proj", D, D)
        ln(p + "norm2")
        lin(p + "mlp.fc1", 4 * D, D)
        lin(p + "mlp.fc2", D, 4 * D)
    ln(pre + "norm")
    if num_classes > 0:
        lin(pre + "head", num_classes, D)
    return sd
