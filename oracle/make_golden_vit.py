"""Gaze-wrapper goldens (TEST INFRASTRUCTURE): runs the UNMODIFIED reference wrappers
``/root/reference/3_Models/backbones/early_fusion_vit.py`` (:103-243) and ``late_fusion_vit.py`` (:118-252) on CPU with
``oracle/timm_stub.py`` standing in for the absent timm (torchvision's VisionTransformer supplies the ViT arithmetic),
and writes ``tests/golden/vit_wrappers.npz``:

  * EarlyFusionViT: logits + ``get_features`` for the 5 fusion modes; the 6-channel patch-embed surgery
    (``duplicate`` / ``average``) applied by the reference constructor to known 3-channel weights;
  * LateFusionViT: logits + ``get_features`` for the 5 fusion modes (eval mode).

Weights are NOT stored (vit_tiny is 22 MB): they are regenerated from the seeds recorded in the fixture by
``oracle.vit.init_vit_state_dict`` (CPU generator, deterministic); a checksum of the weights is stored to detect drift.
Run in the build container only:  python -m oracle.make_golden_vit
"""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import timm_stub  # noqa: E402
from oracle import vit as V  # noqa: E402
from oracle.reference_loader import REFERENCE_ROOT  # noqa: E402
from eyegaze_multimodal_b200.synth import gaze_pair_batch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NAME, B = "vit_tiny_patch16_224", 2
SEED_EARLY6, SEED_EARLY3, SEED_LATE, SEED_IMG, SEED_CLS = 21, 22, 23, 5, 24


def _load(stem):
    path = os.path.join(REFERENCE_ROOT, "3_Models", "backbones", stem + ".py")
    spec = importlib.util.spec_from_file_location("ref_" + stem, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def checksum(sd):
    return np.array([float(sum(v.double().sum() for v in sd.values())), float(sum(v.double().abs().sum() for v in sd.values()))])


def late_classifier(fd, seed=SEED_CLS):
    g = torch.Generator().manual_seed(seed)
    return 0.05 * torch.randn(3, fd, generator=g), 0.05 * torch.randn(3, generator=g)


def main():
    timm_stub.install()
    efv, lfv = _load("early_fusion_vit"), _load("late_fusion_vit")
    a, b = gaze_pair_batch(B, seed=SEED_IMG)
    arrs = {"name": np.array(NAME), "B": np.array(B),
            "seeds": np.array([SEED_EARLY6, SEED_EARLY3, SEED_LATE, SEED_IMG, SEED_CLS])}
    quiet = contextlib.redirect_stdout(io.StringIO())
    with torch.no_grad(), quiet:
        # ---- EarlyFusionViT, 5 fusion modes --------------------------------------------------------------------
        for mode in V.EARLY_MODES:
            cin = 6 if mode == "concat" else 3
            sd = V.init_vit_state_dict(NAME, cin, 3, "backbone.", seed=SEED_EARLY6 if cin == 6 else SEED_EARLY3)
            m = efv.EarlyFusionViT(NAME, num_classes=3, pretrained=False, fusion_mode=mode).eval()
            m.backbone.load_timm_state_dict(sd, "backbone.")
            arrs[f"early::{mode}::logits"] = m(a, b).numpy()
            arrs[f"early::{mode}::features"] = m.get_features(a, b).numpy()
            arrs[f"early::{mode}::checksum"] = checksum(sd)
        # ---- 6-channel surgery applied BY THE REFERENCE CONSTRUCTOR to known 3-channel weights ------------------
        sd3 = V.init_vit_state_dict(NAME, 3, 3, "backbone.", seed=SEED_EARLY3)
        real_create = sys.modules["timm"].create_model

        def create_with_weights(*args, **kw):
            mm = real_create(*args, **kw)
            mm.load_timm_state_dict(sd3, "backbone.")
            return mm
        for strategy in ("duplicate", "average"):
            sys.modules["timm"].create_model = create_with_weights
            efv.timm.create_model = create_with_weights
            try:
                m = efv.EarlyFusionViT(NAME, num_classes=3, pretrained=False, fusion_mode="concat",
                                       weight_init_strategy=strategy).eval()
            finally:
                sys.modules["timm"].create_model = real_create
                efv.timm.create_model = real_create
            w6 = m.backbone.patch_embed.proj.weight.detach()
            arrs[f"surgery::{strategy}::w6_head"] = w6[:4].numpy()            # first 4 output channels, all 6 inputs
            arrs[f"surgery::{strategy}::w6_sum"] = np.array([float(w6.double().sum()), float(w6.double().abs().sum())])
            arrs[f"surgery::{strategy}::bias"] = m.backbone.patch_embed.proj.bias.detach().numpy()
            arrs[f"surgery::{strategy}::logits"] = m(a, b).numpy()
        # ---- LateFusionViT, 5 fusion modes ------------------------------------------------------------------------
        sdl = V.init_vit_state_dict(NAME, 3, 0, "encoder.", seed=SEED_LATE)
        for mode in V.LATE_MODES:
            m = lfv.LateFusionViT(NAME, num_classes=3, pretrained=False, fusion_mode=mode).eval()
            m.encoder.load_timm_state_dict(sdl, "encoder.")
            w, bias = late_classifier(m.fused_dim)
            m.classifier.weight.copy_(w)
            m.classifier.bias.copy_(bias)
            arrs[f"late::{mode}::logits"] = m(a, b).numpy()
            f = m.get_features(a, b)
            arrs[f"late::{mode}::fused"] = f["fused"].numpy()
            arrs[f"late::{mode}::fused_dim"] = np.array(m.fused_dim)
        arrs["late::cls1"] = f["cls1"].numpy()
        arrs["late::cls2"] = f["cls2"].numpy()
        arrs["late::checksum"] = checksum(sdl)
    path = os.path.join(GOLD, "vit_wrappers.npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
