"""CPU-only: on-disk formats either side of the path.

* the ``best_model.pt`` / ``checkpoint-epoch-N.pt`` dictionaries of train_art.py:467-486 round-trip through the drop-in
  DualEEGTransformer with weights saved from the UNMODIFIED reference (tests/golden);
* ``load_pretrained_encoder`` (train_multimodal_fuzzy_fusion.py:285-315: filter the checkpoint by the encoder's own
  keys, update, load) works against the drop-in modules, also from a full multimodal checkpoint with prefixes;
* a timm-keyed 3-channel ImageNet ViT checkpoint initialises the 6-channel early-fusion backbone the way
  early_fusion_vit.py:133-147 does (duplicate / average), through $EGB_VIT_CHECKPOINT_DIR.
No kernel is launched."""
import ast
import os
import warnings

import torch

from conftest import golden_state_dict, load_golden
from oracle import vit as V

NAME = "vit_tiny_patch16_224"


def _eeg_from_golden(name="full"):
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    g = load_golden(f"eeg_model_{name}.npz")
    kw = ast.literal_eval(str(g["kwargs_repr"]))
    return DualEEGTransformer(**kw), kw, golden_state_dict(g)


def test_best_model_pt_round_trip(tmp_path):
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    model, kw, sd = _eeg_from_golden()
    model.load_state_dict(sd, strict=True)                      # weights written by the reference's own module
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    path = tmp_path / "best_model.pt"
    torch.save({"epoch": 7, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "best_f1": 0.61, "config": {"model": kw}}, path)             # train_art.py:467-475
    ck = torch.load(path, map_location="cpu", weights_only=False)
    fresh = DualEEGTransformer(**ck["config"]["model"])
    fresh.load_state_dict(ck["model_state_dict"], strict=True)
    for k, v in sd.items():
        assert torch.equal(fresh.state_dict()[k], v), k
    torch.optim.AdamW(fresh.parameters(), lr=1e-4).load_state_dict(ck["optimizer_state_dict"])
    assert ck["epoch"] == 7 and abs(ck["best_f1"] - 0.61) < 1e-12


def _load_pretrained_encoder(encoder, checkpoint):
    """The logic of train_multimodal_fuzzy_fusion.py:298-314, restated for the test."""
    state_dict = checkpoint["model_state_dict"] if "model_state_dict" in checkpoint else checkpoint
    own = encoder.state_dict()
    pre = {k: v for k, v in state_dict.items() if k in own}
    own.update(pre)
    encoder.load_state_dict(own)
    return len(pre), len(own)


def test_load_pretrained_encoder_flow(tmp_path):
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    eeg, kw, sd = _eeg_from_golden("a1_baseline")
    n, total = _load_pretrained_encoder(eeg, {"model_state_dict": sd})
    assert n == total == len(sd)
    assert all(torch.equal(eeg.state_dict()[k], v) for k, v in sd.items())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gaze = EarlyFusionViT(NAME, num_classes=3, pretrained=False, fusion_mode="concat")
    vsd = V.init_vit_state_dict(NAME, 6, 3, "backbone.", seed=5)
    n, total = _load_pretrained_encoder(gaze, vsd)               # bare state_dict format
    assert n == total
    assert torch.equal(gaze.state_dict()["backbone.blocks.0.attn.qkv.weight"], vsd["backbone.blocks.0.attn.qkv.weight"])
    # a checkpoint of another model (no matching keys) leaves the encoder untouched, as in the reference
    before = {k: v.clone() for k, v in gaze.state_dict().items()}
    n, _ = _load_pretrained_encoder(gaze, {"model_state_dict": {"something.else": torch.zeros(3)}})
    assert n == 0 and all(torch.equal(before[k], v) for k, v in gaze.state_dict().items())


def test_timm_keyed_imagenet_checkpoint_initialises_six_channel_backbone(tmp_path, monkeypatch):
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    timm_sd = V.init_vit_state_dict(NAME, 3, 1000, "", seed=11)   # timm layout: 3-channel patch embed, 1000-way head
    torch.save(timm_sd, tmp_path / (NAME + ".pth"))
    monkeypatch.setenv("EGB_VIT_CHECKPOINT_DIR", str(tmp_path))
    for strategy in ("duplicate", "average"):
        with warnings.catch_warnings():
            warnings.simplefilter("error")                        # the "no checkpoint" warning must NOT fire
            m = EarlyFusionViT(NAME, num_classes=3, pretrained=True, fusion_mode="concat", weight_init_strategy=strategy)
        w = m.backbone.patch_embed.proj.weight.detach()
        w3 = timm_sd["patch_embed.proj.weight"]
        assert w.shape[1] == 6 and torch.equal(w[:, :3], w3)      # early_fusion_vit.py:133-147
        if strategy == "duplicate":
            assert torch.equal(w[:, 3:], w3)
        else:
            assert torch.allclose(w[:, 3:], w3.mean(1, keepdim=True).expand_as(w3))
        assert torch.equal(m.backbone.blocks[3].mlp.fc1.weight.detach(), timm_sd["blocks.3.mlp.fc1.weight"])
        assert m.backbone.head.weight.shape == (3, w3.shape[0])   # the 1000-way head is not carried over
    # other fusion modes keep the 3-channel stem
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m = EarlyFusionViT(NAME, num_classes=3, pretrained=True, fusion_mode="add")
    assert torch.equal(m.backbone.patch_embed.proj.weight.detach(), timm_sd["patch_embed.proj.weight"])
