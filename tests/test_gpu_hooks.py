"""Hook points the reference's analysis code attaches to the model (SURVEY 8b "hook points that must survive", 8f-3),
exercised on the GPU the way 5_Metrics/eeg_metrics.py and 6_Utils/attention_utils.py use them:

  * eeg_metrics.py:195-205  forward hook on ``model.ibs_matrix_generator`` observing (B, 6, 7, C, C);
  * eeg_metrics.py:332-343  forward hook that REPLACES the output (band masking) -> logits change, and equal the oracle
                            evaluated on the masked matrices;
  * eeg_metrics.py:432-453,515-524  forward hook on ``model.cross_attn.cross_attn.dropout``: input[0] = softmax
                            probabilities (B, H, L, L), fired twice per forward (z1->z2, z2->z1);
  * eeg_metrics.py:742-765,841  Grad-CAM: forward + full-backward hooks on ``spectrogram_generator.spec_conv[3]`` with
                            all parameters frozen and the EEG inputs requiring grad;
  * attention_utils.py:196-215  forward + full-backward hooks on ``backbone.blocks[-1]`` of the gaze encoder.
Run on the B200 box:  pytest -m gpu"""
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200 import _lib as L
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import eeg_pair_batch, gaze_pair_batch
from oracle import eeg as O
from oracle import vit as V

DEV = "cuda:0"
CFG = O.EEGConfig(in_channels=8, d_model=64, num_layers=2, num_heads=4, d_ff=128, max_len=96)
B, T = 3, 256


def _model(seed=3):
    sd = O.init_state_dict(CFG, seed)
    m = DualEEGTransformer(**{k: getattr(CFG, k) for k in CFG.__dataclass_fields__})
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


def test_ibs_matrix_hook_observes_and_band_masking_replaces_output(cuda_device):
    m, sd = _model()
    e1, e2 = eeg_pair_batch(B, 8, T, seed=5, coupled=True)
    seen = []
    h = m.ibs_matrix_generator.register_forward_hook(lambda mod, inp, out: seen.append(out.detach().cpu().numpy()))
    with precision("fp32"), torch.no_grad():
        base = m(e1.to(DEV), e2.to(DEV))["logits"].cpu()
    h.remove()
    assert len(seen) == 1 and seen[0].shape == (B, 6, 7, 8, 8)
    mats = O.ibs_connectivity(e1, e2, 256.0, "all")

    def mask_band(mod, inp, out):                   # eeg_metrics.py:334-338: in-place edit + return
        out[:, 3, :, :, :] = 0
        return out
    h = m.ibs_matrix_generator.register_forward_hook(mask_band)
    with precision("fp32"), torch.no_grad():
        masked = m(e1.to(DEV), e2.to(DEV))["logits"].cpu()
    h.remove()
    assert (masked - base).abs().max() > 1e-6       # the replacement reached the tokenizer
    mm = mats.clone()
    mm[:, 3] = 0
    want = O.dual_eeg_forward(sd, e1, e2, CFG, None, ibs_matrices=mm)["logits"]
    want_base = O.dual_eeg_forward(sd, e1, e2, CFG, None, ibs_matrices=mats)["logits"]
    assert (base - want_base).abs().max() <= 1e-4
    assert (masked - want).abs().max() <= 1e-4


def test_cross_attention_probability_hook(cuda_device):
    m, sd = _model()
    e1, e2 = eeg_pair_batch(B, 8, T, seed=6, coupled=True)
    weights = []

    def hook(mod, inp, out):                        # eeg_metrics.py:444-453
        weights.append((inp[0] if isinstance(inp, tuple) else inp).detach().cpu())
    h = m.cross_attn.cross_attn.dropout.register_forward_hook(hook)
    n0 = L.launch_count()
    with precision("fp32"), torch.no_grad():
        hooked = m(e1.to(DEV), e2.to(DEV))["logits"].cpu()
    assert L.launch_count() > n0                    # the hooked forward still runs on this library's kernels
    h.remove()
    with precision("fp32"), torch.no_grad():
        plain = m(e1.to(DEV), e2.to(DEV))["logits"].cpu()
    assert torch.equal(hooked, plain)               # observing does not change the result
    Lseq = 1 + 42 + 8 + T // 16
    assert len(weights) == 2                        # z1 -> z2, then z2 -> z1 (eeg_metrics.py:515-524)
    for w in weights:
        assert w.shape == (B, CFG.num_heads, Lseq, Lseq)
        assert (w.sum(-1) - 1).abs().max() <= 1e-5 and w.min() >= 0
    # against the oracle's softmax(QK^T / sqrt(dk)) on the encoder outputs (art.py:206-208)
    taps = {}
    O.dual_eeg_forward(sd, e1, e2, CFG, None, taps=taps)
    assert (weights[0] - taps["cross_probs"][0]).abs().max() <= 2e-5
    assert (weights[1] - taps["cross_probs"][1]).abs().max() <= 2e-5
    # bf16 mode exports probabilities too (fp32 tensors, rows sum to 1)
    weights.clear()
    h = m.cross_attn.cross_attn.dropout.register_forward_hook(hook)
    with precision("bf16"), torch.no_grad():
        m(e1.to(DEV), e2.to(DEV))
    h.remove()
    assert len(weights) == 2 and (weights[0].float().sum(-1) - 1).abs().max() <= 2e-2


class _GradCAM:
    """eeg_metrics.py:742-765, verbatim structure."""

    def __init__(self, target_layer):
        self.activations, self.gradients = [], []
        self.hooks = [target_layer.register_forward_hook(lambda mod, i, o: self.activations.append(o.detach())),
                      target_layer.register_full_backward_hook(lambda mod, gi, go: self.gradients.append(go[0].detach()))]

    def remove(self):
        for h in self.hooks:
            h.remove()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_gradcam_hooks_on_spec_conv3(cuda_device, mode):
    m, sd = _model()
    e1, e2 = eeg_pair_batch(B, 8, T, seed=7, coupled=True)
    # reference first: the oracle's op chain (det:98-127) in fp32 on the CPU, gradients through the whole model.  The
    # class whose score is back-propagated is the ORACLE's prediction in both passes (a random-init model has margins of
    # a few 1e-2, so a bf16 forward may rank the classes differently; the hook contract is what is under test here).
    sdr = {k: v.clone() for k, v in sd.items()}
    taps = {}
    out = O.dual_eeg_forward(sdr, e1.clone().requires_grad_(True), e2.clone().requires_grad_(True), CFG, None, taps=taps)
    one_hot = F.one_hot(out["logits"].argmax(1), 3).float()
    (out["logits"] * one_hot).sum().backward()
    a1, a2 = taps["spec_conv3"]
    cam = _GradCAM(m.spectrogram_generator.spec_conv[3])
    for p in m.parameters():                        # eeg_metrics.py:861-862
        p.requires_grad = False
    x1 = e1.to(DEV).requires_grad_(True)
    x2 = e2.to(DEV).requires_grad_(True)
    n0 = L.launch_count()
    with precision(mode):
        logits = m(x1, x2)["logits"]
        (logits * one_hot.to(DEV)).sum().backward()                  # eeg_metrics.py:887-889
    launches = L.launch_count() - n0
    cam.remove()
    assert len(cam.activations) == 2 and len(cam.gradients) == 2     # eeg_metrics.py:895
    frames = 1 + T // 64
    shape = (B * 8, 64, 32, frames // 2)
    for t in cam.activations + cam.gradients:
        assert tuple(t.shape) == shape and t.dtype == torch.float32
    tol = 2e-4 if mode == "fp32" else 4e-2
    for got, want in zip(cam.activations, (a1, a2)):                  # forward order: player 1, player 2
        assert (got.cpu() - want.detach()).abs().max() <= tol * want.detach().abs().max(), "activation"
    # backward hooks fire in reverse order: player 2 first (eeg_metrics.py:900-905).  fp32: element-wise; bf16: with three
    # trials through a random-init model a single ReLU unit whose pre-activation sits at ~0 flips under bf16 rounding and
    # changes a trial's whole gradient map, so the bf16 maps are compared as maps (Frobenius-relative + correlation)
    for got, want in zip(cam.gradients, (a2.grad, a1.grad)):
        if mode == "fp32":
            assert (got.cpu() - want).abs().max() <= tol * want.abs().max() + 1e-9, "gradient"
        else:
            g, w = got.cpu().flatten().double(), want.flatten().double()
            assert (g - w).norm() <= 0.6 * w.norm(), "gradient map"
            assert torch.dot(g, w) / (g.norm() * w.norm()) >= 0.8, "gradient map correlation"
    # served by this library's kernels, not by an ATen re-computation of the branch
    assert launches > 50
    for p in m.parameters():
        p.requires_grad = True


def test_gradcam_hooks_on_last_vit_block(cuda_device):
    """attention_utils.py:196-215: activations / gradients of backbone.blocks[-1], (B, 197, D)."""
    warnings.simplefilter("ignore")
    name, heads = "vit_tiny_patch16_224", 3
    sd = V.init_vit_state_dict(name, 6, 3, "backbone.", seed=3)
    m = EarlyFusionViT(name, num_classes=3, pretrained=False, fusion_mode="concat")
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    a, b = gaze_pair_batch(1, seed=2)
    acts, grads = [], []
    layer = m.backbone.blocks[-1]
    h1 = layer.register_forward_hook(lambda mod, i, o: acts.append(o.detach()))
    h2 = layer.register_full_backward_hook(lambda mod, gi, go: grads.append(go[0].detach()))
    try:
        with precision("fp32"):
            logits = m(a.to(DEV), b.to(DEV))
            one_hot = torch.zeros_like(logits)
            one_hot[0, logits.argmax(dim=1).item()] = 1
            logits.backward(gradient=one_hot, retain_graph=True)
    finally:
        h1.remove()
        h2.remove()
    assert len(acts) == 1 and len(grads) == 1
    D = V.VIT_VARIANTS[name][0]
    assert tuple(acts[0].shape) == (1, 197, D) and tuple(grads[0].shape) == (1, 197, D)
    # reference: the oracle's tokens after the last block (before the final LayerNorm) and their gradient
    taps = {}
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    f = V.vit_features(torch.cat([a, b], 1), sdr, "backbone.", heads, taps=taps)
    lg = F.linear(f[:, 0], sdr["backbone.head.weight"], sdr["backbone.head.bias"])
    oh = torch.zeros_like(lg)
    oh[0, lg.argmax(1).item()] = 1
    lg.backward(gradient=oh)
    tok = taps["last_block"]
    assert (acts[0].float().cpu() - tok.detach()).abs().max() <= 2e-4 * tok.detach().abs().max()
    assert (grads[0].float().cpu() - tok.grad).abs().max() <= 2e-3 * tok.grad.abs().max() + 1e-9
