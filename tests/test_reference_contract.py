"""Constructor contract of the drop-in boundary, driven the way the reference drives it.

tests/golden/constructor_contract.json was written by oracle/make_golden_contract.py from the UNMODIFIED reference:
``run_experiments.load_base_config`` + ``create_experiment_config`` for all 13 ``EXPERIMENTS`` and the
``DualEEGTransformer(...)`` call of train_art.py:352-385; ``multimodal_fuzzy_fusion.yaml`` through the constructor calls
of train_multimodal_fuzzy_fusion.py:653-710.  Here the same kwargs go through the OVERLAY files, bound by file path exactly
as the reference's scripts bind their model files (train_art.py:31-44), and the resulting modules must expose the same
``state_dict`` keys / shapes, parameter counts and optional attributes.  CPU only (constructors launch no kernels)."""
import importlib.util
import json
import os
import sys
import warnings

import pytest

from conftest import GOLD, ROOT

OVERLAY = os.path.join(ROOT, "overlay", "3_Models")


def import_module_from_path(module_name, file_path):
    """train_art.py:31-37, verbatim semantics: spec_from_file_location + sys.modules registration."""
    spec = importlib.util.spec_from_file_location(module_name, file_path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[module_name] = module
    spec.loader.exec_module(module)
    return module


@pytest.fixture(scope="module")
def bound():
    saved = {k: sys.modules.get(k) for k in ("dual_eeg_transformer", "early_fusion_vit", "late_fusion_vit",
                                             "fuzzy_gating_fusion", "art")}
    path0 = list(sys.path)
    sys.path.insert(0, os.path.join(OVERLAY, "backbones"))          # train_art.py:26
    mods = {
        "det": import_module_from_path("dual_eeg_transformer", os.path.join(OVERLAY, "backbones", "dual_eeg_transformer.py")),
        "efv": import_module_from_path("early_fusion_vit", os.path.join(OVERLAY, "backbones", "early_fusion_vit.py")),
        "lfv": import_module_from_path("late_fusion_vit", os.path.join(OVERLAY, "backbones", "late_fusion_vit.py")),
        "fgf": import_module_from_path("fuzzy_gating_fusion", os.path.join(OVERLAY, "fusion", "fuzzy_gating_fusion.py")),
    }
    yield mods
    sys.path[:] = path0
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


@pytest.fixture(scope="module")
def contract():
    with open(os.path.join(GOLD, "constructor_contract.json")) as f:
        return json.load(f)


def _check(model, want):
    sd = model.state_dict()
    assert [[k, list(v.shape)] for k, v in sd.items()] == want["state_dict"]
    assert sum(p.numel() for p in model.parameters()) == want["n_params"]
    for flag, present in want["flags"].items():
        assert hasattr(model, flag) == present, flag


def test_all_ablation_experiments_build_through_the_overlay(bound, contract):
    assert len(contract["experiments"]) == 13
    for name, case in contract["experiments"].items():
        model = bound["det"].DualEEGTransformer(**case["kwargs"])
        _check(model, case)
        # the loss methods the loop calls when the experiment enables them (train_art.py:193-214)
        for meth in ("compute_symmetry_loss", "compute_ibs_alignment_loss", "compute_ibs_contrastive_loss"):
            assert callable(getattr(model, meth)), (name, meth)


def test_multimodal_config_builds_through_the_overlay(bound, contract):
    from eyegaze_multimodal_b200.multimodal import MultimodalFusionModel
    mm = contract["multimodal"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                # pretrained=True without network: random init
        gaze = bound["efv"].EarlyFusionViT(**mm["gaze_kwargs"])
    eeg = bound["det"].DualEEGTransformer(**mm["eeg_kwargs"])
    fusion = bound["fgf"].FuzzyGatingFusion(**mm["fusion_kwargs"])
    _check(eeg, mm["eeg"])
    assert [[k, list(v.shape)] for k, v in fusion.state_dict().items()] == mm["fusion"]["state_dict"]
    assert sum(p.numel() for p in gaze.parameters()) == 86_390_787      # 4_Experiments/experiments_list.md:62
    assert gaze.backbone.patch_embed.proj.in_channels == 6 and hasattr(gaze.backbone, "blocks")
    model = MultimodalFusionModel(gaze_encoder=gaze, eeg_encoder=eeg, fusion_module=fusion,
                                  freeze_gaze=mm["freeze"]["gaze"], freeze_eeg=mm["freeze"]["eeg"])
    keys = list(model.state_dict())
    assert any(k.startswith("gaze_encoder.backbone.blocks.11.mlp.fc2.") for k in keys)
    assert any(k.startswith("eeg_encoder.encoder.layers.5.") for k in keys) and "fusion.beta" in keys
    # the two learning-rate groups of train_multimodal_fuzzy_fusion.py:727-736 partition the parameters
    enc = [p for n, p in model.named_parameters() if not n.startswith("fusion.")]
    fus = [p for n, p in model.named_parameters() if n.startswith("fusion.")]
    assert len(fus) == 9 and len(enc) + len(fus) == len(list(model.parameters()))


def test_error_conventions(bound):
    with pytest.raises(ValueError):
        bound["fgf"].FuzzyGatingFusion(mode="nope")
    with pytest.raises(ValueError):
        bound["efv"].EarlyFusionViT(pretrained=False, fusion_mode="full")
    with pytest.raises(ValueError):
        bound["lfv"].LateFusionViT(pretrained=False, fusion_mode="subtract_abs")
    with pytest.raises(AssertionError):
        bound["det"].DualEEGTransformer(in_channels=8, d_model=30, num_heads=4)     # art.py:172
    m = bound["det"].DualEEGTransformer(in_channels=8, ibs_feature_type="bogus", max_len=64)
    assert m.ibs_matrix_generator.num_features == 7                                 # unknown type falls through to "all"


def test_fixture_is_current_with_the_reference(contract):
    """Where the reference tree exists (build container), the recorded kwargs are re-derived live."""
    from oracle.reference_loader import REFERENCE_ROOT, available
    if not available():
        pytest.skip("reference tree not present (GPU box)")
    from oracle.make_golden_contract import eeg_kwargs_train_art
    spec = importlib.util.spec_from_file_location("ref_run_experiments", os.path.join(REFERENCE_ROOT, "run_experiments.py"))
    rx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rx)
    base = rx.load_base_config()
    assert set(rx.EXPERIMENTS) == set(contract["experiments"])
    for name, exp in rx.EXPERIMENTS.items():
        assert eeg_kwargs_train_art(rx.create_experiment_config(base, exp, name)) == contract["experiments"][name]["kwargs"]
