"""Input side of the path on the GPU: the per-sample normalisation the reference's datasets run
on the host, applied to a whole device batch by two HBM-bound kernels.

* ``normalize_eeg_windows`` -- ``DualEEGDataset._preprocess_eeg`` (common average reference + per-channel z-score,
  ``1_Data/processed/dual_eeg_dataset.py:158-166``) or the plain window z-score of ``:196-198``; the host ships raw windows.
* ``normalize_images_u8`` -- ``transforms.ToTensor()`` + ``transforms.Normalize(mean, std)``
  (``1_Data/processed/multimodal_dataset.py:73-83``) from uint8 HWC batches: a quarter of the fp32 H2D bytes.
"""
import ctypes as C

import torch

from . import _lib as L
from . import torch_ops as TO

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def normalize_eeg_windows(x: torch.Tensor, enable_preprocessing: bool = True) -> torch.Tensor:
    """x: (B, C, T) or (C, T) fp32 CUDA windows -> normalised windows of the same shape."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise TypeError("normalize_eeg_windows expects an fp32 CUDA tensor (no CPU path)")
    squeeze = x.dim() == 2
    xb = (x.unsqueeze(0) if squeeze else x).contiguous()
    if xb.dim() != 3:
        raise ValueError("expected (B, C, T) or (C, T)")
    out = torch.empty_like(xb)
    TO.call("eeg_window_normalize", xb, out, xb.shape[0], xb.shape[1], xb.shape[2],
           0 if enable_preprocessing else 1)
    return out.squeeze(0) if squeeze else out


def normalize_images_u8(u8_hwc: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """u8_hwc: (B, H, W, 3) uint8 CUDA batch (PIL / numpy layout) -> (B, 3, H, W) fp32, ToTensor + Normalize."""
    if not u8_hwc.is_cuda or u8_hwc.dtype != torch.uint8:
        raise TypeError("normalize_images_u8 expects a uint8 CUDA tensor (no CPU path)")
    if u8_hwc.dim() != 4 or u8_hwc.shape[-1] != 3:
        raise ValueError("expected (B, H, W, 3)")
    x = u8_hwc.contiguous()
    B, H, W, _ = x.shape
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
    TO.call("image_u8_normalize", x, out, B, H, W, [float(v) for v in mean], [float(v) for v in std])
    return out
