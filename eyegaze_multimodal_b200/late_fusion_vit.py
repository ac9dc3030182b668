"""LateFusionViT -- drop-in for ``3_Models/backbones/late_fusion_vit.py`` (cited ``lfv:<line>``).

Siamese ViT encoder (``.encoder``, timm keys, ``num_classes=0``) -> CLS-feature fusion -> Dropout -> Linear.
Both heat-maps go through the shared encoder as ONE stacked 2B batch (the reference runs it twice, lfv:214-215).
"""
from typing import Literal

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from . import vit as _vit

FUSION_MODES = Literal['concat', 'add', 'subtract', 'multiply', 'full']


class LateFusionViT(nn.Module):
    def __init__(self, model_name: str = 'vit_base_patch16_224', num_classes: int = 3, pretrained: bool = True,
                 fusion_mode: FUSION_MODES = 'full', dropout: float = 0.1):
        super().__init__()
        self.model_name = model_name
        self.num_classes = num_classes
        self.fusion_mode = fusion_mode
        valid_modes = ['concat', 'add', 'subtract', 'multiply', 'full']
        if fusion_mode not in valid_modes:
            raise ValueError(f"fusion_mode must be one of {valid_modes}, got '{fusion_mode}'")
        self.encoder = _vit.create_model(model_name, pretrained=pretrained, num_classes=0)
        self.embed_dim = self.encoder.num_features
        if fusion_mode == 'concat':
            self.fused_dim = 2 * self.embed_dim
        elif fusion_mode in ['add', 'subtract', 'multiply']:
            self.fused_dim = self.embed_dim
        else:
            self.fused_dim = 4 * self.embed_dim
        self.dropout = nn.Dropout(p=dropout)
        self.classifier = nn.Linear(self.fused_dim, num_classes)

    def _fuse_features(self, cls1: torch.Tensor, cls2: torch.Tensor) -> torch.Tensor:
        """lfv:148-178 on (B, D) fp32 features (a few kB: plain tensor ops)."""
        if self.fusion_mode == 'concat':
            return torch.cat([cls1, cls2], dim=1)
        if self.fusion_mode == 'add':
            return cls1 + cls2
        if self.fusion_mode == 'subtract':
            return cls1 - cls2
        if self.fusion_mode == 'multiply':
            return cls1 * cls2
        return torch.cat([cls1, cls2, cls1 - cls2, cls1 * cls2], dim=1)

    def _cls_pair(self, x1, x2):
        tok = self.encoder._encode(self.encoder.embed(x1, x2, "single_pair"))      # (2B, 1+N, D)
        cls = self.encoder.forward_head(tok, pre_logits=True)                     # (2B, D) fp32
        B = x1.shape[0]
        return cls[:B], cls[B:]

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        cls1, cls2 = self._cls_pair(x1, x2)
        fused = self._fuse_features(cls1, cls2)
        p = self.dropout.p if self.training else 0.0
        return ops.linear(fused.contiguous(), self.classifier.weight, self.classifier.bias, out_f32=True) if p == 0.0 \
            else ops.linear(self.dropout(fused).contiguous(), self.classifier.weight, self.classifier.bias, out_f32=True)

    def get_features(self, x1: torch.Tensor, x2: torch.Tensor) -> dict:
        cls1, cls2 = self._cls_pair(x1, x2)
        return {'cls1': cls1, 'cls2': cls2, 'fused': self._fuse_features(cls1, cls2)}


def create_late_fusion_vit(model_name: str = 'vit_base_patch16_224', num_classes: int = 3, pretrained: bool = True,
                           fusion_mode: str = 'full', **kwargs) -> LateFusionViT:
    return LateFusionViT(model_name=model_name, num_classes=num_classes, pretrained=pretrained, fusion_mode=fusion_mode,
                         **kwargs)
