"""MultimodalFusionModel + its loss -- the glue module the reference defines inside its training script
(4_Experiments/scripts/train_multimodal_fuzzy_fusion.py:106-179, loss at :436-460)."""
import os
import warnings
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops


class MultimodalFusionModel(nn.Module):
    def __init__(self, gaze_encoder: nn.Module, eeg_encoder: nn.Module, fusion_module: nn.Module,
                 freeze_gaze: bool = False, freeze_eeg: bool = False):
        super().__init__()
        self.gaze_encoder = gaze_encoder
        self.eeg_encoder = eeg_encoder
        self.fusion = fusion_module
        if freeze_gaze:
            for p in self.gaze_encoder.parameters():
                p.requires_grad = False
        if freeze_eeg:
            for p in self.eeg_encoder.parameters():
                p.requires_grad = False

    # The two encoders are independent until the fusion head (train_multimodal_fuzzy_fusion.py:162-166), forward
    # and backward.  On CUDA the EEG branch is enqueued on a side stream: its latency-bound kernels (FFT / connectivity
    # / spectrogram prologue, small-K GEMMs) fill the gaps the ViT's big GEMMs leave, and autograd replays each
    # branch's backward on the stream its forward ran on, so the backward overlaps the same way.
    concurrent_branches = os.environ.get("EGB_CONCURRENT_BRANCHES", "1") != "0"

    # autograd notes that parameters first touched on the side stream accumulate their gradient there; that is the
    # intended schedule (the engine inserts the cross-stream waits itself), so the advisory is silenced
    warnings.filterwarnings("ignore", message="The AccumulateGrad node's stream does not match")

    def forward(self, img1: torch.Tensor, img2: torch.Tensor, eeg1: torch.Tensor, eeg2: torch.Tensor,
                labels: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if self.concurrent_branches and eeg1.is_cuda:
            cur = torch.cuda.current_stream(eeg1.device)
            side = getattr(self, "_side_stream", None)
            if side is None or side.device != eeg1.device:
                side = torch.cuda.Stream(device=eeg1.device)
                object.__setattr__(self, "_side_stream", side)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                eeg_logits = self.eeg_encoder(eeg1, eeg2, labels)['logits']
            img_logits = self.gaze_encoder(img1, img2)
            cur.wait_stream(side)
            if not torch.cuda.is_current_stream_capturing():
                eeg_logits.record_stream(cur)
        else:
            img_logits = self.gaze_encoder(img1, img2)
            eeg_logits = self.eeg_encoder(eeg1, eeg2, labels)['logits']
        fused_logits, alpha, aux_info = self.fusion(img_logits, eeg_logits)
        return {'fused_logits': fused_logits, 'img_logits': img_logits, 'eeg_logits': eeg_logits, 'alpha': alpha,
                'aux_info': aux_info}


def multimodal_loss(model: MultimodalFusionModel, outputs: Dict, labels: torch.Tensor, lambda_aux_img: float = 0.3,
                    lambda_aux_eeg: float = 0.3, lambda_reg: float = 0.1, t_min: float = 0.5,
                    t_max: float = 5.0) -> torch.Tensor:
    """CE(fused) + 0.3 CE(img/T_img) + 0.3 CE(eeg/T_eeg) + 0.1 temperature hinge; temperatures are the detached
    copies from aux_info (train_multimodal_fuzzy_fusion.py:440-460)."""
    t = outputs['aux_info']['temperatures']
    loss = ops.cross_entropy(outputs['fused_logits'], labels)
    loss = loss + lambda_aux_img * ops.cross_entropy(outputs['img_logits'] / t['img'], labels)
    loss = loss + lambda_aux_eeg * ops.cross_entropy(outputs['eeg_logits'] / t['eeg'], labels)
    return loss + lambda_reg * model.fusion.compute_temperature_regularization(t_min, t_max).squeeze()
