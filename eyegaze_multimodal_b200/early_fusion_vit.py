"""EarlyFusionViT -- drop-in for ``3_Models/backbones/early_fusion_vit.py`` (cited ``efv:<line>``).

Same constructor, ``fusion_mode`` / ``weight_init_strategy`` semantics, ``.backbone`` attribute (timm key names) and
``forward`` / ``get_features``.  The input fusion (efv:163-196) is folded into the patch-extraction kernel, so the
6-channel concatenation (308 MB at batch 256) is never written.
"""
from typing import Literal

import torch
import torch.nn as nn

from . import vit as _vit

FUSION_MODES = Literal['concat', 'add', 'subtract', 'subtract_abs', 'multiply']


class EarlyFusionViT(nn.Module):
    def __init__(self, model_name: str = 'vit_base_patch16_224', num_classes: int = 3, pretrained: bool = True,
                 img_size: int = 224, fusion_mode: FUSION_MODES = 'concat',
                 weight_init_strategy: Literal['duplicate', 'average'] = 'duplicate'):
        super().__init__()
        self.model_name = model_name
        self.num_classes = num_classes
        self.fusion_mode = fusion_mode
        self.weight_init_strategy = weight_init_strategy
        valid_modes = ['concat', 'add', 'subtract', 'subtract_abs', 'multiply']
        if fusion_mode not in valid_modes:
            raise ValueError(f"fusion_mode must be one of {valid_modes}, got '{fusion_mode}'")
        self.backbone = _vit.create_model(model_name, pretrained=pretrained, num_classes=num_classes, img_size=img_size)
        if fusion_mode == 'concat':
            self._modify_patch_embed_for_6_channels()

    def _modify_patch_embed_for_6_channels(self):
        """efv:103-147: widen the 3-channel patch projection to 6 channels, seeding both halves from the original."""
        old = self.backbone.patch_embed.proj
        new = nn.Conv2d(6, old.out_channels, kernel_size=old.kernel_size, stride=old.stride, padding=old.padding,
                        bias=old.bias is not None)
        with torch.no_grad():
            w = old.weight.data.clone()
            new.weight[:, 0:3] = w
            if self.weight_init_strategy == 'duplicate':
                new.weight[:, 3:6] = w
            elif self.weight_init_strategy == 'average':
                new.weight[:, 3:6] = w.mean(dim=1, keepdim=True).expand_as(w)
            if old.bias is not None:
                new.bias.data = old.bias.data.clone()
        self.backbone.patch_embed.proj = new

    def _fuse_inputs(self, img_a: torch.Tensor, img_b: torch.Tensor) -> torch.Tensor:
        """Materialised fusion (efv:163-196) for callers that want the fused image itself (analysis only)."""
        if self.fusion_mode == 'concat':
            return torch.cat([img_a, img_b], dim=1)
        if self.fusion_mode == 'add':
            return (img_a + img_b) / 2.0
        if self.fusion_mode == 'subtract':
            return (img_a - img_b) / 2.0
        if self.fusion_mode == 'subtract_abs':
            return torch.abs(img_a - img_b)
        prod = img_a * img_b
        B, C, H, W = prod.shape
        flat = prod.view(B, C, -1)
        flat = (flat - flat.mean(dim=2, keepdim=True)) / (flat.std(dim=2, keepdim=True) + 1e-6)
        return flat.view(B, C, H, W)

    def _tokens(self, img_a, img_b):
        return self.backbone.forward_fused_pair(img_a, img_b, self.fusion_mode)

    def forward(self, img_a: torch.Tensor, img_b: torch.Tensor) -> torch.Tensor:
        return self.backbone.forward_head(self._tokens(img_a, img_b))

    def get_features(self, img_a: torch.Tensor, img_b: torch.Tensor) -> torch.Tensor:
        return self.backbone.forward_head(self._tokens(img_a, img_b), pre_logits=True)


def create_early_fusion_vit(model_name: str = 'vit_base_patch16_224', num_classes: int = 3, pretrained: bool = True,
                            fusion_mode: str = 'concat', **kwargs) -> EarlyFusionViT:
    return EarlyFusionViT(model_name=model_name, num_classes=num_classes, pretrained=pretrained, fusion_mode=fusion_mode,
                          **kwargs)
