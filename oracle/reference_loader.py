"""Imports the UNMODIFIED reference modules by file path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Mirrors how the reference's own scripts bind them (4_Experiments/scripts/train_art.py:31-44,
train_multimodal_fuzzy_fusion.py:62-88): ``importlib.util.spec_from_file_location`` + registration in
``sys.modules`` under the bare file stem, with ``3_Models/backbones`` on ``sys.path`` so that
``dual_eeg_transformer`` finds ``art``.  Only usable where the reference tree exists (this container);
nothing that runs on the GPU box may call it.
"""
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("EGB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "3_Models", "backbones", "dual_eeg_transformer.py"))


def _load(stem: str, rel: str, alias: str):
    path = os.path.join(REFERENCE_ROOT, rel)
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with the reference's art, dual_eeg_transformer and fuzzy_gating_fusion modules.

    The modules are registered under ``ref_*`` aliases (plus a temporary bare ``art`` entry while
    ``dual_eeg_transformer`` resolves its import) so they never shadow the drop-in modules of the
    same stem that the product ships.
    """
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in ("art", "hf_config")}
    try:
        sys.modules.pop("hf_config", None)
        art = _load("art", "3_Models/backbones/art.py", "ref_art")
        sys.modules["art"] = art
        det = _load("dual_eeg_transformer", "3_Models/backbones/dual_eeg_transformer.py", "ref_dual_eeg_transformer")
        fgf = _load("fuzzy_gating_fusion", "3_Models/fusion/fuzzy_gating_fusion.py", "ref_fuzzy_gating_fusion")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    class NS:
        pass

    ns = NS()
    ns.art, ns.det, ns.fgf = art, det, fgf
    return ns
