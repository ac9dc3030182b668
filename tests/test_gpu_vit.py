"""GPU parity of the gaze encoders and the full multimodal model against the CPU oracle (oracle/vit.py is
cross-checked against torchvision in test_oracle_vit.py).  Run on the B200 box:  pytest -m gpu"""
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    from eyegaze_multimodal_b200.late_fusion_vit import LateFusionViT
    from eyegaze_multimodal_b200.multimodal import MultimodalFusionModel, multimodal_loss
    from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import eeg_pair_batch, gaze_pair_batch
from oracle import eeg as O
from oracle import fuzzy as FZ
from oracle import vit as V

DEV = "cuda:0"
NAME, HEADS = "vit_tiny_patch16_224", 3


def _early(mode, seed=0):
    warnings.simplefilter("ignore")
    cin = 6 if mode == "concat" else 3
    sd = V.init_vit_state_dict(NAME, cin, 3, "backbone.", seed=seed)
    m = EarlyFusionViT(NAME, num_classes=3, pretrained=False, fusion_mode=mode)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("mode", ["concat", "add", "subtract", "subtract_abs", "multiply"])
def test_early_fusion_forward_fp32_bf16(cuda_device, mode):
    m, sd = _early(mode)
    a, b = gaze_pair_batch(2, seed=1)
    want = V.early_fusion_forward(sd, a, b, HEADS, mode)
    with precision("fp32"), torch.no_grad():
        got = m(a.to(DEV), b.to(DEV))
        feat = m.get_features(a.to(DEV), b.to(DEV))
    assert (got.cpu() - want).abs().max() <= 1e-4
    assert (feat.cpu() - V.early_fusion_features(sd, a, b, HEADS, mode)).abs().max() <= 2e-4
    with precision("bf16"), torch.no_grad():
        gb = m(a.to(DEV), b.to(DEV))
    assert (gb.float().cpu() - want).abs().max() / want.abs().max() <= 2e-2


def test_early_fusion_backward_fp32(cuda_device):
    m, sd = _early("concat", seed=3)
    a, b = gaze_pair_batch(2, seed=2)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    labels = torch.tensor([0, 2])
    F.cross_entropy(V.early_fusion_forward(sdr, a, b, HEADS, "concat"), labels).backward()
    with precision("fp32"):
        F.cross_entropy(m(a.to(DEV), b.to(DEV)), labels.to(DEV)).backward()
    for k, p in m.named_parameters():
        r = sdr[k].grad
        e = (p.grad.cpu() - r).abs().max().item()
        assert e <= 5e-3 * r.abs().max().item() + 2e-6, f"{k}: {e:.3e} vs max {r.abs().max().item():.3e}"


def test_early_fusion_backward_bf16(cuda_device):
    """bf16 mode: EVERY backbone parameter must receive a gradient (the dtype conversions stay on the autograd
    graph) and it must agree with the fp32 CPU oracle to bf16 accuracy (relative to the tensor's max)."""
    m, sd = _early("concat", seed=3)
    a, b = gaze_pair_batch(2, seed=2)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    labels = torch.tensor([0, 2])
    F.cross_entropy(V.early_fusion_forward(sdr, a, b, HEADS, "concat"), labels).backward()
    with precision("bf16"):
        F.cross_entropy(m(a.to(DEV), b.to(DEV)), labels.to(DEV)).backward()
    for k, p in m.named_parameters():
        assert p.grad is not None, f"{k} received no gradient in bf16 mode"
        r = sdr[k].grad
        e = (p.grad.cpu() - r).abs().max().item()
        assert e <= 8e-2 * r.abs().max().item() + 1e-5, f"{k}: {e:.3e} vs max {r.abs().max().item():.3e}"


def test_patch_embed_surgery(cuda_device):
    warnings.simplefilter("ignore")
    torch.manual_seed(0)
    m = EarlyFusionViT(NAME, pretrained=False, fusion_mode="concat", weight_init_strategy="duplicate")
    w = m.backbone.patch_embed.proj.weight
    assert w.shape[1] == 6 and torch.equal(w[:, :3], w[:, 3:])
    m = EarlyFusionViT(NAME, pretrained=False, fusion_mode="concat", weight_init_strategy="average")
    w = m.backbone.patch_embed.proj.weight
    assert torch.allclose(w[:, 3:], w[:, :3].mean(1, keepdim=True).expand_as(w[:, :3]))
    with pytest.raises(ValueError):
        EarlyFusionViT(NAME, pretrained=False, fusion_mode="full")


@pytest.mark.parametrize("mode", ["full", "concat", "multiply"])
def test_late_fusion(cuda_device, mode):
    warnings.simplefilter("ignore")
    sd = V.init_vit_state_dict(NAME, 3, 0, "encoder.", seed=4)
    D = sd["encoder.norm.weight"].shape[0]
    fd = {"concat": 2 * D, "add": D, "subtract": D, "multiply": D, "full": 4 * D}[mode]
    g = torch.Generator().manual_seed(5)
    sd["classifier.weight"] = 0.05 * torch.randn(3, fd, generator=g)
    sd["classifier.bias"] = 0.05 * torch.randn(3, generator=g)
    m = LateFusionViT(NAME, pretrained=False, fusion_mode=mode)
    assert m.fused_dim == fd
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x1, x2 = gaze_pair_batch(2, seed=6)
    want = V.late_fusion_forward(sd, x1, x2, HEADS, mode)
    with precision("fp32"), torch.no_grad():
        got = m(x1.to(DEV), x2.to(DEV))
        feats = m.get_features(x1.to(DEV), x2.to(DEV))
    assert (got.cpu() - want).abs().max() <= 1e-4
    assert feats["fused"].shape == (2, fd) and feats["cls1"].shape == (2, D)


def test_multimodal_end_to_end(cuda_device):
    """cfg 2 composition at reduced size: EarlyFusionViT(concat) + DualEEGTransformer + FuzzyGatingFusion and the
    4-term loss, forward and parameter gradients against the oracle."""
    warnings.simplefilter("ignore")
    cfg = O.EEGConfig(in_channels=8, d_model=64, num_layers=2, num_heads=4, d_ff=128, max_len=96)
    esd = O.init_state_dict(cfg, 7)
    vsd = V.init_vit_state_dict(NAME, 6, 3, "backbone.", seed=8)
    fsd = FZ.init_params()
    eeg = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    eeg.load_state_dict(esd, strict=True)
    gaze = EarlyFusionViT(NAME, pretrained=False, fusion_mode="concat")
    gaze.load_state_dict(vsd, strict=True)
    model = MultimodalFusionModel(gaze, eeg, FuzzyGatingFusion(3, "full")).to(DEV).eval()
    B = 3
    e1, e2 = eeg_pair_batch(B, 8, 256, seed=9, coupled=True)
    a, b = gaze_pair_batch(B, seed=10)
    labels = torch.tensor([0, 1, 2])
    # oracle
    er = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in esd.items()}
    vr = {k: v.clone().requires_grad_(True) for k, v in vsd.items()}
    fr = {k: v.clone().requires_grad_(k != "c_reliable") for k, v in fsd.items()}
    il = V.early_fusion_forward(vr, a, b, HEADS, "concat")
    el = O.dual_eeg_forward(er, e1, e2, cfg, labels)["logits"]
    fused, alpha, aux = FZ.fuzzy_forward(fr, il, el, "full")
    loss_ref = FZ.multimodal_loss(fused, il, el, aux, FZ.temperature_regularization(fr), labels)
    loss_ref.backward()
    with precision("fp32"):
        out = model(a.to(DEV), b.to(DEV), e1.to(DEV), e2.to(DEV), labels.to(DEV))
        loss = multimodal_loss(model, out, labels.to(DEV))
        loss.backward()
    assert (out["fused_logits"].detach().cpu() - fused.detach()).abs().max() <= 1e-4
    assert (out["alpha"].detach().cpu() - alpha.detach()).abs().max() <= 1e-5
    assert abs(loss.item() - loss_ref.item()) <= 1e-4
    assert set(out["aux_info"]) == {"temperatures", "entropies", "membership", "firing_strengths", "consequents", "fuzz_params"}
    assert not out["aux_info"]["entropies"]["img"].requires_grad
    for name, ref in (("tau_img", fr), ("beta", fr), ("c_unreliable_eeg", fr)):
        gp = getattr(model.fusion, name).grad.cpu()
        assert (gp - ref[name].grad).abs().max() <= 5e-3 * ref[name].grad.abs().max() + 1e-6, name
    for k in ("backbone.head.weight", "backbone.blocks.0.attn.qkv.weight", "backbone.patch_embed.proj.weight"):
        gp = dict(model.gaze_encoder.named_parameters())[k].grad.cpu()
        assert (gp - vr[k].grad).abs().max() <= 5e-3 * vr[k].grad.abs().max() + 1e-6, k
    for k in ("classifier.3.weight", "temporal_conv.convs.0.weight", "encoder.layers.1.ffn.linear1.weight"):
        gp = dict(model.eeg_encoder.named_parameters())[k].grad.cpu()
        assert (gp - er[k].grad).abs().max() <= 5e-3 * er[k].grad.abs().max() + 1e-6, k
    # bf16 mode: the whole model still trains (all parameters that feed the loss get finite gradients)
    model.zero_grad(set_to_none=True)
    with precision("bf16"):
        o16 = model(a.to(DEV), b.to(DEV), e1.to(DEV), e2.to(DEV), labels.to(DEV))
        multimodal_loss(model, o16, labels.to(DEV)).backward()
    for k, p in model.named_parameters():
        if "ibs_classifier" in k:
            continue                      # auxiliary head: not part of the multimodal loss
        assert p.grad is not None and torch.isfinite(p.grad).all(), f"{k}: no / non-finite gradient in bf16 mode"
    for k in ("backbone.blocks.0.attn.qkv.weight", "backbone.patch_embed.proj.weight"):
        gp = dict(model.gaze_encoder.named_parameters())[k].grad.cpu()
        assert (gp - vr[k].grad).abs().max() <= 0.15 * vr[k].grad.abs().max() + 1e-4, k
    model.zero_grad(set_to_none=True)
    # frozen encoders receive no gradient (train_multimodal_fuzzy_fusion.py:129-137)
    frozen = MultimodalFusionModel(gaze, eeg, FuzzyGatingFusion(3, "full"), freeze_gaze=True, freeze_eeg=True).to(DEV)
    frozen.zero_grad()
    with precision("fp32"):
        o2 = frozen(a.to(DEV), b.to(DEV), e1.to(DEV), e2.to(DEV))
        multimodal_loss(frozen, o2, labels.to(DEV)).backward()
    assert all(p.grad is None or p.grad.abs().max() == 0 for p in frozen.gaze_encoder.parameters())
    assert frozen.fusion.tau_img.grad is not None
