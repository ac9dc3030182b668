"""ViT oracle (oracle/vit.py): timm is absent, so the restatement is cross-checked against torchvision's
VisionTransformer and Hugging Face's ViTForImageClassification by weight remapping, and against the parameter count the
reference documents. CPU only."""
import numpy as np
import pytest
import torch

from oracle import vit as V


def _tv_to_timm(tv_sd, depth, pre):
    sd = {pre + "cls_token": tv_sd["class_token"], pre + "pos_embed": tv_sd["encoder.pos_embedding"],
          pre + "patch_embed.proj.weight": tv_sd["conv_proj.weight"], pre + "patch_embed.proj.bias": tv_sd["conv_proj.bias"],
          pre + "norm.weight": tv_sd["encoder.ln.weight"], pre + "norm.bias": tv_sd["encoder.ln.bias"]}
    for i in range(depth):
        s, d = f"encoder.layers.encoder_layer_{i}.", f"{pre}blocks.{i}."
        for a, b in (("ln_1", "norm1"), ("ln_2", "norm2"), ("self_attention.out_proj", "attn.proj"),
                     ("mlp.0", "mlp.fc1"), ("mlp.3", "mlp.fc2")):
            sd[d + b + ".weight"], sd[d + b + ".bias"] = tv_sd[s + a + ".weight"], tv_sd[s + a + ".bias"]
        sd[d + "attn.qkv.weight"], sd[d + "attn.qkv.bias"] = tv_sd[s + "self_attention.in_proj_weight"], tv_sd[s + "self_attention.in_proj_bias"]
    if "heads.head.weight" in tv_sd:
        sd[pre + "head.weight"], sd[pre + "head.bias"] = tv_sd["heads.head.weight"], tv_sd["heads.head.bias"]
    return sd


def test_vit_restatement_matches_torchvision():
    tvm = pytest.importorskip("torchvision.models.vision_transformer")
    torch.manual_seed(0)
    m = tvm.VisionTransformer(image_size=32, patch_size=16, num_layers=2, num_heads=4, hidden_dim=64, mlp_dim=256,
                              num_classes=3).eval()
    m.conv_proj = torch.nn.Conv2d(6, 64, 16, 16)
    for p in m.parameters():
        torch.nn.init.normal_(p, std=0.05)
    sd = _tv_to_timm({k: v.detach() for k, v in m.state_dict().items()}, 2, "backbone.")
    a, b = torch.randn(3, 3, 32, 32), torch.randn(3, 3, 32, 32)
    with torch.no_grad():
        want = m(torch.cat([a, b], 1))
    got = V.early_fusion_forward(sd, a, b, heads=4, mode="concat")
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=2e-5, rtol=1e-4)


def test_documented_parameter_count():
    """4_Experiments/experiments_list.md:62 -- EarlyFusionViT ViT-B/16, 6 channels, 3 classes."""
    sd = V.init_vit_state_dict("vit_base_patch16_224", 6, 3, "backbone.")
    assert sum(v.numel() for v in sd.values()) == 86_390_787


def test_wrapper_logic():
    a, b = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    assert V.fuse_inputs(a, b, "concat").shape == (2, 6, 8, 8)
    assert torch.allclose(V.fuse_inputs(a, b, "add"), (a + b) / 2)
    assert torch.allclose(V.fuse_inputs(a, b, "subtract_abs"), (a - b).abs())
    m = V.fuse_inputs(a, b, "multiply").view(2, 3, -1)
    assert torch.allclose(m.mean(2), torch.zeros(2, 3), atol=1e-5) and torch.allclose(m.std(2), torch.ones(2, 3), atol=1e-3)
    with pytest.raises(ValueError):
        V.fuse_inputs(a, b, "full")
    c1, c2 = torch.randn(2, 5), torch.randn(2, 5)
    dims = {"concat": 10, "add": 5, "subtract": 5, "multiply": 5, "full": 20}  # late_fusion_vit.py:304-310
    for mode, d in dims.items():
        assert V.fuse_features(c1, c2, mode).shape == (2, d)
    w3 = torch.randn(4, 3, 16, 16)
    w6 = V.widen_patch_embed(w3, "duplicate")
    assert torch.equal(w6[:, :3], w3) and torch.equal(w6[:, 3:], w3)
    w6 = V.widen_patch_embed(w3, "average")
    assert torch.allclose(w6[:, 3:], w3.mean(1, keepdim=True).expand_as(w3))


def _hf_to_timm(hf_sd, depth, pre):
    """Hugging Face ViTForImageClassification keys -> timm keys (q / k / v rows stacked into the fused qkv)."""
    e = "vit.embeddings."
    sd = {pre + "cls_token": hf_sd[e + "cls_token"], pre + "pos_embed": hf_sd[e + "position_embeddings"],
          pre + "patch_embed.proj.weight": hf_sd[e + "patch_embeddings.projection.weight"],
          pre + "patch_embed.proj.bias": hf_sd[e + "patch_embeddings.projection.bias"],
          pre + "norm.weight": hf_sd["vit.layernorm.weight"], pre + "norm.bias": hf_sd["vit.layernorm.bias"],
          pre + "head.weight": hf_sd["classifier.weight"], pre + "head.bias": hf_sd["classifier.bias"]}
    for i in range(depth):
        s, d = f"vit.encoder.layer.{i}.", f"{pre}blocks.{i}."
        for a, b in (("layernorm_before", "norm1"), ("layernorm_after", "norm2"), ("attention.output.dense", "attn.proj"),
                     ("intermediate.dense", "mlp.fc1"), ("output.dense", "mlp.fc2")):
            sd[d + b + ".weight"], sd[d + b + ".bias"] = hf_sd[s + a + ".weight"], hf_sd[s + a + ".bias"]
        for part in ("weight", "bias"):
            sd[d + "attn.qkv." + part] = torch.cat([hf_sd[s + f"attention.attention.{n}.{part}"] for n in ("query", "key", "value")])
    return sd


def test_vit_restatement_matches_huggingface_vit():
    """A second, independent implementation of the published ViT (transformers' ViTForImageClassification: separate
    q/k/v projections, its own attention and embedding code) agrees with the restatement after key remapping."""
    tr = pytest.importorskip("transformers")
    cfg = tr.ViTConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256, image_size=32,
                       patch_size=16, num_channels=6, num_labels=3, hidden_act="gelu", layer_norm_eps=1e-6,
                       hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, qkv_bias=True)
    torch.manual_seed(1)
    m = tr.ViTForImageClassification(cfg).eval()
    for p in m.parameters():
        torch.nn.init.normal_(p, std=0.05)
    sd = _hf_to_timm({k: v.detach() for k, v in m.state_dict().items()}, 2, "backbone.")
    a, b = torch.randn(3, 3, 32, 32), torch.randn(3, 3, 32, 32)
    with torch.no_grad():
        want = m(pixel_values=torch.cat([a, b], 1)).logits
    got = V.early_fusion_forward(sd, a, b, heads=4, mode="concat")
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=2e-5, rtol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# Wrapper logic PINNED to the unmodified reference: tests/golden/vit_wrappers.npz was written by oracle/make_golden_vit.py,
# which runs /root/reference/3_Models/backbones/{early,late}_fusion_vit.py unchanged with oracle/timm_stub.py (torchvision
# ViT arithmetic behind timm's attribute surface) in place of the absent timm.  Weights are regenerated from the seeds.
# ---------------------------------------------------------------------------------------------------------------------
from conftest import load_golden  # noqa: E402
from eyegaze_multimodal_b200.synth import gaze_pair_batch  # noqa: E402


def _wrapper_golden():
    g = load_golden("vit_wrappers.npz")
    s6, s3, sl, simg, scls = (int(x) for x in g["seeds"])
    return g, str(g["name"]), int(g["B"]), s6, s3, sl, simg, scls


def _checksum(sd):
    return np.array([float(sum(v.double().sum() for v in sd.values())), float(sum(v.double().abs().sum() for v in sd.values()))])


@pytest.mark.parametrize("mode", V.EARLY_MODES)
def test_early_wrapper_matches_reference_golden(mode):
    g, name, B, s6, s3, _sl, simg, _ = _wrapper_golden()
    heads = V.VIT_VARIANTS[name][2]
    cin = 6 if mode == "concat" else 3
    sd = V.init_vit_state_dict(name, cin, 3, "backbone.", seed=s6 if cin == 6 else s3)
    np.testing.assert_allclose(_checksum(sd), g[f"early::{mode}::checksum"], rtol=1e-9)     # same regenerated weights
    a, b = gaze_pair_batch(B, seed=simg)
    with torch.no_grad():
        logits = V.early_fusion_forward(sd, a, b, heads, mode)
        feats = V.early_fusion_features(sd, a, b, heads, mode)
    np.testing.assert_allclose(logits.numpy(), g[f"early::{mode}::logits"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(feats.numpy(), g[f"early::{mode}::features"], atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("strategy", ["duplicate", "average"])
def test_patch_embed_surgery_matches_reference_golden(strategy):
    g, name, B, _s6, s3, _sl, simg, _ = _wrapper_golden()
    heads = V.VIT_VARIANTS[name][2]
    sd = V.init_vit_state_dict(name, 3, 3, "backbone.", seed=s3)
    w6 = V.widen_patch_embed(sd["backbone.patch_embed.proj.weight"], strategy)
    np.testing.assert_allclose(w6[:4].numpy(), g[f"surgery::{strategy}::w6_head"], atol=1e-7)
    np.testing.assert_allclose([float(w6.double().sum()), float(w6.double().abs().sum())], g[f"surgery::{strategy}::w6_sum"], rtol=1e-6)
    np.testing.assert_allclose(sd["backbone.patch_embed.proj.bias"].numpy(), g[f"surgery::{strategy}::bias"], atol=0)
    sd["backbone.patch_embed.proj.weight"] = w6
    a, b = gaze_pair_batch(B, seed=simg)
    with torch.no_grad():
        logits = V.early_fusion_forward(sd, a, b, heads, "concat")
    np.testing.assert_allclose(logits.numpy(), g[f"surgery::{strategy}::logits"], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("mode", V.LATE_MODES)
def test_late_wrapper_matches_reference_golden(mode):
    g, name, B, _s6, _s3, sl, simg, scls = _wrapper_golden()
    heads = V.VIT_VARIANTS[name][2]
    sd = V.init_vit_state_dict(name, 3, 0, "encoder.", seed=sl)
    np.testing.assert_allclose(_checksum(sd), g["late::checksum"], rtol=1e-9)
    fd = int(g[f"late::{mode}::fused_dim"])
    gen = torch.Generator().manual_seed(scls)
    sd["classifier.weight"] = 0.05 * torch.randn(3, fd, generator=gen)
    sd["classifier.bias"] = 0.05 * torch.randn(3, generator=gen)
    x1, x2 = gaze_pair_batch(B, seed=simg)
    with torch.no_grad():
        logits = V.late_fusion_forward(sd, x1, x2, heads, mode)
        c1 = V.vit_features(x1, sd, "encoder.", heads)[:, 0]
        c2 = V.vit_features(x2, sd, "encoder.", heads)[:, 0]
    np.testing.assert_allclose(logits.numpy(), g[f"late::{mode}::logits"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(V.fuse_features(c1, c2, mode).numpy(), g[f"late::{mode}::fused"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(c1.numpy(), g["late::cls1"], atol=5e-5, rtol=1e-4)
