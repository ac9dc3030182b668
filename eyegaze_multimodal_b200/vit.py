"""VisionTransformer with timm's ``vit_*_patch16_224`` arithmetic, attribute names and ``state_dict`` keys
(``cls_token, pos_embed, patch_embed.proj.*, blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*, norm.*,
head.*``), running on this package's sm_100a kernels.

The reference obtains this model from ``timm.create_model`` (early_fusion_vit.py:86-91, late_fusion_vit.py:106-110);
timm is a third-party dependency that is not vendored with the reference, so the definition below follows timm's
published ViT: pre-LN blocks, LayerNorm eps 1e-6, fused qkv Linear with bias, softmax((q*hd^-0.5) k^T) v, exact-erf
GELU MLP (ratio 4), learned pos_embed (1, 1+N, D), CLS pooling, no dropout / drop-path at the default settings.

Per block the kernels are:  LN -> qkv GEMM -> fused attention -> proj GEMM (+bias +residual epilogue) -> LN ->
fc1 GEMM (+bias +GELU epilogue, pre-activation kept for the backward) -> fc2 GEMM (+bias +residual epilogue).
"""
import math
import os
import warnings

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .art import LayerNorm, _has_hooks
from .precision import compute_code

# timm registry entries reachable from the reference's configs: (embed_dim, depth, heads)
VIT_VARIANTS = {
    "vit_tiny_patch16_224": (192, 12, 3),
    "vit_small_patch16_224": (384, 12, 6),
    "vit_base_patch16_224": (768, 12, 12),
    "vit_large_patch16_224": (1024, 24, 16),
}


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.Identity()


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def context(self, x):
        qkv = ops.linear(x, self.qkv.weight, self.qkv.bias)
        return ops.attention_packed(qkv, self.num_heads, scale=self.scale)

    def forward(self, x):
        x = ops.cast(x, compute_code())
        return ops.linear(self.context(x), self.proj.weight, self.proj.bias)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def fused(self, x, residual=None):
        return ops.mlp2(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, L.ACT_GELU, residual=residual)

    def forward(self, x):
        return self.fused(ops.cast(x, compute_code()))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = ops.cast(x, compute_code())
        # pre-norm residual blocks: the normalisation hands back x as the residual operand, so the two gradients
        # that meet at x are summed inside the LayerNorm backward kernel
        h, x = ops.layernorm_residual(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        ctx = self.attn.context(h)
        x = ops.linear(ctx, self.attn.proj.weight, self.attn.proj.bias, residual=x)
        h, x = ops.layernorm_residual(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.mlp.fused(h, residual=x)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_prefix_tokens = 1
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n + 1, embed_dim) * .02)
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self._init_weights()

    def _init_weights(self):
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    # -- token embedding ---------------------------------------------------------------------------------
    def embed(self, img_a, img_b=None, mode="single"):
        """Fused input-fusion + patch embedding + cls + pos: (B,3,H,W)[x2] -> (B or 2B, 1+N, D)."""
        pe = self.patch_embed.proj
        return ops.vit_embed(img_a, img_b, pe.weight, pe.bias, self.cls_token, self.pos_embed, ops.PATCH_MODES[mode],
                             compute_code(), self.patch_embed.patch_size[0])

    def _tokens_from_tensor(self, x):
        cin = self.patch_embed.proj.in_channels
        if x.shape[1] != cin:
            raise RuntimeError("expected %d input channels, got %d" % (cin, x.shape[1]))
        if cin == 6:
            return self.embed(x[:, :3], x[:, 3:], "concat")     # channel-sliced views, no copy
        if cin == 3:
            return self.embed(x, None, "single")
        raise NotImplementedError("patch embedding supports 3 or 6 input channels")

    def _encode(self, t):
        t = self.blocks(t)
        return self.norm(t)

    def forward_features(self, x):
        return self._encode(self._tokens_from_tensor(x))

    def forward_head(self, x, pre_logits: bool = False):
        cls = ops.cast(x[:, 0].contiguous(), L.F32)              # CLS pooling (timm global_pool='token')
        if pre_logits or isinstance(self.head, nn.Identity):
            return cls
        return ops.linear(cls, self.head.weight, self.head.bias, out_f32=True)

    def forward(self, x):
        return self.forward_head(self.forward_features(x))

    def forward_fused_pair(self, img_a, img_b, mode):
        """Tokens of the early-fused pair without materialising the fused image."""
        return self._encode(self.embed(img_a, img_b, mode))


def _load_local_pretrained(model: VisionTransformer, model_name: str) -> bool:
    path = os.environ.get("EGB_VIT_CHECKPOINT_DIR")
    if not path:
        return False
    f = os.path.join(path, model_name + ".pth")
    if not os.path.isfile(f):
        return False
    sd = torch.load(f, map_location="cpu")
    sd = sd.get("state_dict", sd)
    own = model.state_dict()
    keep = {k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}
    model.load_state_dict(keep, strict=False)
    return True


def create_model(model_name: str, pretrained: bool = False, num_classes: int = 1000, img_size: int = 224,
                 in_chans: int = 3, **kwargs) -> VisionTransformer:
    """Stand-in for ``timm.create_model`` for the ViT variants the reference's configs name."""
    if model_name not in VIT_VARIANTS:
        raise ValueError("unknown model %r: the B200 build provides %s" % (model_name, sorted(VIT_VARIANTS)))
    dim, depth, heads = VIT_VARIANTS[model_name]
    m = VisionTransformer(img_size=img_size, patch_size=16, in_chans=in_chans, num_classes=num_classes, embed_dim=dim,
                          depth=depth, num_heads=heads)
    if pretrained and not _load_local_pretrained(m, model_name):
        warnings.warn("pretrained=True: no network here and no checkpoint under $EGB_VIT_CHECKPOINT_DIR/%s.pth -- "
                      "the ViT keeps its random initialisation" % model_name)
    return m
