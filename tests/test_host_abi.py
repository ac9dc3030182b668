"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/eyegaze_b200.h declares,
the drop-in modules keep the reference's constructor signatures and state_dict keys, the overlay files import, and the
product path refuses to run without CUDA (no CPU fallback).  No kernel is launched here."""
import ctypes
import importlib.util
import inspect
import os
import re
import sys

import pytest
import torch

from conftest import ROOT, golden_state_dict, load_golden
from eyegaze_multimodal_b200 import _lib


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "eyegaze_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(egb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_lib.LIB_PATH), "build the library first: python -m eyegaze_multimodal_b200.csrc.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [s for s in declared if s not in _lib.EXPORTED_SYMBOLS]
    assert not unbound, f"declared in the header but not bound by _lib.py: {unbound}"
    lib.egb_version.restype = ctypes.c_int
    assert lib.egb_version() >= 1


def test_constructor_signatures_match_reference_contract():
    """constructor argument names and defaults are the drop-in contract."""
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    from eyegaze_multimodal_b200.late_fusion_vit import LateFusionViT
    sig = inspect.signature(DualEEGTransformer.__init__)
    want = dict(in_channels=62, num_classes=3, d_model=256, num_layers=6, num_heads=8, d_ff=1024, dropout=0.1, max_len=2048,
                conv_kernel_size=25, conv_stride=4, conv_layers=2, sampling_rate=256, use_spectrogram=True, spec_n_fft=128,
                spec_hop_length=64, spec_freq_bins=64, use_robust_ibs=True, use_ibs=True, use_cross_attention=True,
                ibs_instance_norm=True, ibs_feature_type="all")
    got = {k: v.default for k, v in sig.parameters.items() if k != "self"}
    assert got == want
    assert {k: v.default for k, v in inspect.signature(FuzzyGatingFusion.__init__).parameters.items() if k != "self"} == \
        dict(num_classes=3, mode="full", eps_temp=0.1, eps_log=1e-8, eps_div=1e-8)
    assert list(inspect.signature(EarlyFusionViT.__init__).parameters)[1:] == \
        ["model_name", "num_classes", "pretrained", "img_size", "fusion_mode", "weight_init_strategy"]
    assert list(inspect.signature(LateFusionViT.__init__).parameters)[1:] == \
        ["model_name", "num_classes", "pretrained", "fusion_mode", "dropout"]
    with pytest.raises(ValueError):
        FuzzyGatingFusion(mode="nope")
    with pytest.raises(ValueError):
        LateFusionViT("vit_tiny_patch16_224", pretrained=False, fusion_mode="nope")


@pytest.mark.parametrize("name", ["full", "a1_baseline", "scalar_ibs", "phase_noin_nocross"])
def test_state_dict_keys_equal_the_reference_checkpoints(name):
    """Golden state_dicts were saved from the UNMODIFIED reference: strict loading proves key/shape identity, and
    ablation-disabled sub-modules must be absent attributes (5_Metrics/eeg_metrics.py uses hasattr as feature test)."""
    import ast
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    g = load_golden(f"eeg_model_{name}.npz")
    kw = ast.literal_eval(str(g["kwargs_repr"]))
    m = DualEEGTransformer(**kw)
    sd = golden_state_dict(g)
    m.load_state_dict(sd, strict=True)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert hasattr(m, "spectrogram_generator") == kw.get("use_spectrogram", True)
    assert hasattr(m, "cross_attn") == kw.get("use_cross_attention", True)
    assert hasattr(m, "ibs_matrix_generator") == (kw.get("use_ibs", True) and kw.get("use_robust_ibs", True))


def test_fuzzy_state_dict_and_vit_keys():
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    from oracle import vit as V
    f = FuzzyGatingFusion()
    assert set(f.state_dict()) == {"tau_img", "tau_eeg", "c_reliable", "c_unreliable_img", "c_unreliable_eeg",
                                   "log_sigma_reliable_img", "log_sigma_reliable_eeg", "log_sigma_unreliable_img",
                                   "log_sigma_unreliable_eeg", "beta"}
    assert abs(float(f.temp_img) - 1.5) < 1e-6 and abs(float(f.temp_eeg) - 1.0) < 1e-6   # reference self-test values
    assert float(f.compute_temperature_regularization()) == 0.0
    m = EarlyFusionViT("vit_base_patch16_224", num_classes=3, pretrained=False, fusion_mode="concat")
    assert sum(p.numel() for p in m.parameters()) == 86_390_787                           # experiments_list.md:62
    want = V.init_vit_state_dict("vit_base_patch16_224", 6, 3, "backbone.")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in want.items()}


def test_overlay_modules_import_by_file_path_like_the_reference_scripts():
    """train_art.py:31-44 binds the model files with spec_from_file_location under the bare file stem."""
    base = os.path.join(ROOT, "overlay", "3_Models")
    sys.path.insert(0, os.path.join(base, "backbones"))
    try:
        for rel, stem, names in [("backbones/art.py", "art", ["TransformerEncoder", "MultiHeadAttention", "PositionalEmbedding"]),
                                 ("backbones/dual_eeg_transformer.py", "dual_eeg_transformer", ["DualEEGTransformer"]),
                                 ("backbones/early_fusion_vit.py", "early_fusion_vit", ["EarlyFusionViT", "create_early_fusion_vit"]),
                                 ("backbones/late_fusion_vit.py", "late_fusion_vit", ["LateFusionViT", "create_late_fusion_vit"]),
                                 ("fusion/fuzzy_gating_fusion.py", "fuzzy_gating_fusion", ["FuzzyGatingFusion", "inverse_softplus"])]:
            spec = importlib.util.spec_from_file_location(stem, os.path.join(base, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[stem] = mod
            spec.loader.exec_module(mod)
            for n in names:
                assert hasattr(mod, n), f"{rel} does not export {n}"
    finally:
        sys.path.pop(0)
        for stem in ("art", "dual_eeg_transformer", "early_fusion_vit", "late_fusion_vit", "fuzzy_gating_fusion"):
            sys.modules.pop(stem, None)


def test_no_cpu_fallback():
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    m = DualEEGTransformer(in_channels=4, d_model=32, num_layers=1, num_heads=4, d_ff=64, max_len=64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 4, 128), torch.randn(1, 4, 128))
    with pytest.raises(RuntimeError, match="CUDA|CPU"):
        FuzzyGatingFusion()(torch.randn(2, 3), torch.randn(2, 3))
    # nothing in the product package imports the oracle
    pkg = os.path.join(ROOT, "eyegaze_multimodal_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(pkg, fn)).read(), flags=re.M), fn


# ---------------------------------------------------------------------------------------------------------------------
# torch.library registration of the C ABI (north star: "a thin C-ABI torch.library extension")
# ---------------------------------------------------------------------------------------------------------------------
def test_every_c_entry_point_is_a_registered_torch_op():
    import torch
    from eyegaze_multimodal_b200 import torch_ops as T
    host_only = {"egb_prof_enable", "egb_prof_read", "egb_prof_dump", "egb_debug_attention_timing", "egb_debug_gemm_timing",
                 "egb_seed_epoch_enable"}
    want = {n[4:] for n in _lib._SIGNATURES if n not in host_only}
    assert set(T.OP_NAMES) == want and len(want) >= 50
    ns = torch.ops.eyegaze_b200
    for name in want:
        op = getattr(ns, name).default
        s = str(op._schema)
        assert s.startswith("eyegaze_b200::" + name + "(") and s.endswith("-> ()"), s
    # pointers became (optional, mutable-declared) tensors; descriptors are flattened into their fields
    s = str(ns.layernorm_fwd.default._schema)
    assert s.count("Tensor(") == 6 and "float a9" in s
    s = str(ns.gemm.default._schema)
    assert "a0_a_ptr" in s and "a0_c_colsum" in s and "a0_dropout_seed" in s
    s = str(ns.attention_bwd.default._schema)
    assert "a0_dq_colsum" in s and "a0_kv_shift" in s


def test_torch_ops_have_meta_kernels_and_no_cpu_kernel():
    import pytest
    import torch
    ns = torch.ops.eyegaze_b200
    from eyegaze_multimodal_b200 import torch_ops  # noqa: F401  (registers the ops)
    m = torch.empty(4, 8, device="meta")
    v = torch.empty(8, device="meta")
    r = torch.empty(4, device="meta")
    assert ns.layernorm_fwd(m, v, v, m, r, r, 0, 4, 8, 1e-5) is None          # fake-tensor propagation by name
    x = torch.zeros(4, 8)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ns.layernorm_fwd(x, torch.ones(8), torch.zeros(8), x.clone(), torch.zeros(4), torch.zeros(4), 0, 4, 8, 1e-5)


def test_host_side_never_calls_ctypes_directly():
    """ops.py / optim.py / inputs.py reach the kernels through torch.ops.eyegaze_b200 only."""
    import re
    pkg = os.path.join(ROOT, "eyegaze_multimodal_b200")
    for f in ("ops.py", "optim.py", "inputs.py", "art.py", "dual_eeg_transformer.py", "vit.py", "parallel.py", "graphs.py"):
        src = open(os.path.join(pkg, f)).read()
        calls = re.findall(r"L\.call\(\"(egb_\w+)\"", src)
        assert set(calls) <= {"egb_seed_epoch_enable"}, (f, calls)      # host-side allocation helper, not a kernel
        assert "C.byref(" not in src.replace('L.call("egb_seed_epoch_enable", C.byref(out))', ""), f
