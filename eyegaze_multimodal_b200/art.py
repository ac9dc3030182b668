"""Transformer primitives of the EEG encoder -- drop-in for the classes the hot path imports from the
reference's ``3_Models/backbones/art.py`` (lines 55-328): same class names, constructor signatures,
sub-module names and ``state_dict`` keys; forward passes run on the sm_100a kernels of this package.

Fusion map (reference op chain -> kernels here)
  q/k/v projections (art.py:203-205)            -> one packed GEMM (N = 3d) with bias epilogue
  QK^T/sqrt(dk) . softmax . dropout . @V (206-211) -> one fused attention kernel (probabilities never in HBM)
  out_proj + dropout + residual (213, 293)      -> GEMM epilogue (bias, Philox-style dropout mask, residual add)
  LayerNorm (293/295/328)                       -> warp-per-row LN kernel with saved mean / rstd
  linear1 + ReLU + dropout + linear2 + dropout + residual (272, 295) -> two GEMMs with fused epilogues
"""
import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .precision import compute_code


def _has_hooks(*mods) -> bool:
    for m in mods:
        if m is None:
            continue
        if m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or getattr(m, "_backward_pre_hooks", None):
            return True
    return False


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm whose forward is the library's LN kernel (parameters / state_dict keys unchanged)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.elementwise_affine or self.bias is None:
            raise NotImplementedError("LayerNorm without affine parameters is not on the hot path")
        return ops.layernorm(ops.cast(x, compute_code()), self.weight, self.bias, self.eps)


class PositionalEmbedding(nn.Module):
    """art.py:55-126.  mode='learned' (nn.Embedding) is what DualEEGTransformer uses; 'sinusoidal' keeps its buffer."""

    def __init__(self, max_len: int, d_model: int, mode: str = 'sinusoidal') -> None:
        super().__init__()
        if mode not in {'sinusoidal', 'learned'}:
            raise ValueError(f'Unsupported pos_mode: {mode}')
        self.mode = mode
        self.d_model = d_model
        if self.mode == 'learned':
            self.pos_embed = nn.Embedding(max_len, d_model)
        else:
            pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
            div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
            pe = torch.zeros(max_len, d_model)
            pe[:, 0::2] = torch.sin(pos * div)
            pe[:, 1::2] = torch.cos(pos * div)
            self.register_buffer('pe', pe.unsqueeze(0))

    def table(self) -> torch.Tensor:
        return self.pos_embed.weight if self.mode == 'learned' else self.pe[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # standalone use: x + pos[:T]; the model's fast path folds this add into the sequence-assembly kernel
        T = x.size(1)
        tab = self.table()
        if T > tab.shape[0]:
            raise IndexError('index out of range in self')
        return x + tab[:T].to(x.dtype)


class MultiHeadAttention(nn.Module):
    """art.py:128-213."""

    def __init__(self, d_model: int, num_heads: int, dropout: float = 0.0) -> None:
        super().__init__()
        assert d_model % num_heads == 0, 'd_model must be divisible by num_heads'
        self.d_k = d_model // num_heads
        self.num_heads = num_heads
        self.d_model = d_model
        self.q_proj = nn.Linear(d_model, d_model)
        self.k_proj = nn.Linear(d_model, d_model)
        self.v_proj = nn.Linear(d_model, d_model)
        self.out_proj = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)

    # -- building blocks shared with the fused callers --------------------------------------------------
    def _p_attn(self) -> float:
        return self.dropout.p if self.training else 0.0

    def _fire_dropout_hooks(self, probs: torch.Tensor) -> None:
        """Analysis hooks on ``.dropout`` expect input[0] = softmax probabilities (B,H,Lq,Lk)
        (5_Metrics/eeg_metrics.py:432-453); the fused kernel exports them on request."""
        with torch.no_grad():
            self.dropout(probs)

    def context_packed(self, x: torch.Tensor, kv_shift: int = 0, hook_split: int = 0) -> torch.Tensor:
        """softmax(QK^T)V for q, k, v all projected from ``x`` [S, L, d]; kv_shift pairs batch s with s+kv_shift."""
        qkv = ops.linear_packed(x, [self.q_proj.weight, self.k_proj.weight, self.v_proj.weight],
                                [self.q_proj.bias, self.k_proj.bias, self.v_proj.bias])
        want = _has_hooks(self.dropout)
        out = ops.attention_packed(qkv, self.num_heads, kv_shift=kv_shift, p=self._p_attn(), want_probs=want)
        if want:
            out, probs = out
            if hook_split:
                for part in probs.split(hook_split, dim=0):
                    self._fire_dropout_hooks(part)
            else:
                self._fire_dropout_hooks(probs)
        return out

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is not used by the fusion classifier hot path (art.py:207-208 is only "
                                      "reached by the unused ART decoder)")
        code = compute_code()
        self_attn = (q is k and k is v) or (_alias(q, k) and _alias(q, v))
        kv_same = (k is v) or _alias(k, v)
        q = ops.cast(q, code)
        if self_attn:
            ctx = self.context_packed(q)
        else:
            k = ops.cast(k, code)
            qh = ops.linear(q, self.q_proj.weight, self.q_proj.bias)
            if kv_same:
                kv = ops.linear_packed(k, [self.k_proj.weight, self.v_proj.weight], [self.k_proj.bias, self.v_proj.bias])
                kh, vh = kv[..., :self.d_model], kv[..., self.d_model:]
            else:
                kh = ops.linear(k, self.k_proj.weight, self.k_proj.bias)
                vh = ops.linear(ops.cast(v, code), self.v_proj.weight, self.v_proj.bias)
            want = _has_hooks(self.dropout)
            ctx = ops.attention(qh, kh, vh, self.num_heads, p=self._p_attn(), want_probs=want)
            if want:
                ctx, probs = ctx
                self._fire_dropout_hooks(probs)
        return ops.linear(ctx, self.out_proj.weight, self.out_proj.bias)


def _alias(a, b) -> bool:
    return (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype
            and a.requires_grad == b.requires_grad)


class FeedForward(nn.Module):
    """art.py:215-272: dropout(linear2(dropout(relu(linear1(x)))))."""

    def __init__(self, d_model: int, d_ff: int, dropout: float = 0.0) -> None:
        super().__init__()
        self.linear1 = nn.Linear(d_model, d_ff)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ff, d_model)

    def fused(self, x, residual=None, p_after: float = 0.0):
        """``p_after``: a second dropout the caller applies to the FFN output (the block's drop2); two independent
        masks in sequence equal one mask with keep probability (1-p)(1-p_after), so both fold into one epilogue."""
        p = self.dropout.p if self.training else 0.0
        p_out = 1.0 - (1.0 - p) * (1.0 - (p_after if self.training else 0.0))
        return ops.mlp2(x, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias, L.ACT_RELU,
                        p_mid=p, p_out=p_out, residual=residual)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fused(ops.cast(x, compute_code()))


class TransformerEncoderBlock(nn.Module):
    """art.py:274-296 (post-LN)."""

    def __init__(self, d_model: int, num_heads: int, d_ff: int, dropout: float = 0.0, attn_dropout: float = 0.0) -> None:
        super().__init__()
        self.mha = MultiHeadAttention(d_model, num_heads, dropout=attn_dropout)
        self.drop1 = nn.Dropout(dropout)
        self.ln1 = LayerNorm(d_model, eps=1e-05)
        self.ffn = FeedForward(d_model, d_ff, dropout=dropout)
        self.drop2 = nn.Dropout(dropout)
        self.ln2 = LayerNorm(d_model, eps=1e-05)

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is not used by the fusion classifier hot path")
        x = ops.cast(x, compute_code())
        if _has_hooks(self.mha, self.drop1, self.ffn, self.drop2, self.mha.out_proj, self.ffn.linear1, self.ffn.linear2):
            # reference-structured path so module-level hooks observe the same tensors as upstream
            h = self.mha(x, x, x)
            x = self.ln1(x + self.drop1(h))
            h = self.ffn(x)
            return self.ln2(x + self.drop2(h))
        ctx = self.mha.context_packed(x)
        p1 = self.drop1.p if self.training else 0.0
        y = ops.linear(ctx, self.mha.out_proj.weight, self.mha.out_proj.bias, residual=x, p=p1)   # x + drop1(out_proj(ctx))
        x = self.ln1(y)
        return self.ln2(self.ffn.fused(x, residual=x, p_after=self.drop2.p))                    # x + drop2(ffn(x))


class TransformerEncoder(nn.Module):
    """art.py:298-328."""

    def __init__(self, d_model: int, num_layers: int, num_heads: int, d_ff: int, dropout: float = 0.0,
                 attn_dropout: float = 0.0) -> None:
        super().__init__()
        self.layers = nn.ModuleList([
            TransformerEncoderBlock(d_model=d_model, num_heads=num_heads, d_ff=d_ff, dropout=dropout,
                                    attn_dropout=attn_dropout) for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=1e-05)

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        for layer in self.layers:
            x = layer(x, attn_mask=attn_mask)
        return self.norm(x)
