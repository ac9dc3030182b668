"""Constructor-contract fixture (TEST INFRASTRUCTURE): drives the UNMODIFIED reference exactly as its entry points do and
records what a drop-in must reproduce.

  * ``/root/reference/run_experiments.py``: ``load_base_config()`` (4_Experiments/configs/dual_eeg_transformer.yaml) and
    ``create_experiment_config`` (:242-275) for every entry of ``EXPERIMENTS`` (:47-240);
  * the ``DualEEGTransformer(...)`` call of ``4_Experiments/scripts/train_art.py:352-385`` on each resulting config;
  * ``4_Experiments/configs/multimodal_fuzzy_fusion.yaml`` through the constructor calls of
    ``train_multimodal_fuzzy_fusion.py:653-710`` (EEG encoder + fusion module built from the reference; the gaze encoder's
    constructor arguments are recorded, its timm backbone is not available here).

For every case the fixture (tests/golden/constructor_contract.json) holds the constructor kwargs, the reference model's
``state_dict`` keys + shapes, parameter count and the optional-attribute flags the analysis code tests with ``hasattr``
(5_Metrics/eeg_metrics.py:202,437,834).  Run in the build container only:  python -m oracle.make_golden_contract
"""
import importlib.util
import json
import os
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import REFERENCE_ROOT, load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FLAGS = ("spectrogram_generator", "ibs_matrix_generator", "ibs_tokenizer", "ibs_generator", "ibs_classifier", "cross_attn")


def eeg_kwargs_train_art(config):
    """The keyword mapping of train_art.py:352-385 (config dict -> DualEEGTransformer arguments)."""
    ab = config.get("ablation", {})
    return dict(
        in_channels=config["model"]["in_channels"], num_classes=config["model"]["num_labels"],
        d_model=config["model"]["d_model"], num_layers=config["model"]["num_layers"],
        num_heads=config["model"]["num_heads"], d_ff=config["model"]["d_ff"], dropout=config["training"]["dropout"],
        max_len=config["data"]["window_size"] // 4, conv_kernel_size=config["model"]["conv_kernel_size"],
        conv_stride=config["model"]["conv_stride"], conv_layers=config["model"]["conv_layers"],
        sampling_rate=config["data"]["sampling_rate"], use_spectrogram=ab.get("use_spectrogram", True),
        spec_n_fft=config["model"].get("spec_n_fft", 128), spec_hop_length=config["model"].get("spec_hop_length", 64),
        spec_freq_bins=config["model"].get("spec_freq_bins", 64), use_robust_ibs=(ab.get("ibs_mode", "robust") == "robust"),
        use_ibs=ab.get("use_ibs", True), use_cross_attention=ab.get("use_cross_attention", True),
        ibs_instance_norm=ab.get("ibs_instance_norm", True), ibs_feature_type=ab.get("ibs_feature_type", "all"))


def eeg_kwargs_multimodal(config):
    """train_multimodal_fuzzy_fusion.py:670-688."""
    e = config["eeg_encoder"]
    return dict(
        in_channels=e["in_channels"], num_classes=config["model"]["num_classes"], d_model=e["d_model"],
        num_layers=e["num_layers"], num_heads=e["num_heads"], d_ff=e["d_ff"], dropout=e.get("dropout", 0.1),
        max_len=config["data"]["window_size"] // 4, conv_kernel_size=e.get("conv_kernel_size", 25),
        conv_stride=e.get("conv_stride", 4), conv_layers=e.get("conv_layers", 2),
        sampling_rate=config["data"].get("sampling_rate", 256), use_spectrogram=e.get("use_spectrogram", True),
        use_robust_ibs=e.get("use_robust_ibs", True), use_ibs=e.get("use_ibs", True),
        use_cross_attention=e.get("use_cross_attention", True))


def gaze_kwargs_multimodal(config):
    """train_multimodal_fuzzy_fusion.py:653-660."""
    g = config["gaze_encoder"]
    return dict(model_name=g["model_name"], num_classes=config["model"]["num_classes"], pretrained=g.get("pretrained", True),
                img_size=config["data"].get("image_size", 224), fusion_mode=g.get("fusion_mode", "concat"),
                weight_init_strategy=g.get("weight_init_strategy", "duplicate"))


def describe(model):
    sd = model.state_dict()
    return {"state_dict": [[k, list(v.shape)] for k, v in sd.items()],
            "n_params": sum(p.numel() for p in model.parameters()),
            "flags": {f: hasattr(model, f) for f in FLAGS}}


def main():
    ref = load_reference()
    spec = importlib.util.spec_from_file_location("ref_run_experiments", os.path.join(REFERENCE_ROOT, "run_experiments.py"))
    rx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rx)
    base = rx.load_base_config()
    out = {"experiments": {}, "multimodal": {}}
    for name, exp in rx.EXPERIMENTS.items():
        cfg = rx.create_experiment_config(base, exp, name)
        kw = eeg_kwargs_train_art(cfg)
        m = ref.det.DualEEGTransformer(**kw)
        out["experiments"][name] = {"kwargs": kw, "training": {k: cfg["training"][k] for k in
                                    ("use_sym_loss", "use_ibs_loss", "use_ibs_cls_loss", "use_ibs_contrastive",
                                     "lambda_sym", "lambda_ibs", "lambda_ibs_cls", "lambda_ibs_contrastive")},
                                    **describe(m)}
        print(name, out["experiments"][name]["n_params"], out["experiments"][name]["flags"])
    with open(os.path.join(REFERENCE_ROOT, "4_Experiments", "configs", "multimodal_fuzzy_fusion.yaml"), encoding="utf-8") as f:
        mm = yaml.safe_load(f)
    kw = eeg_kwargs_multimodal(mm)
    out["multimodal"]["eeg_kwargs"] = kw
    out["multimodal"]["eeg"] = describe(ref.det.DualEEGTransformer(**kw))
    out["multimodal"]["gaze_kwargs"] = gaze_kwargs_multimodal(mm)
    fz = ref.fgf.FuzzyGatingFusion(num_classes=mm["model"]["num_classes"], mode=mm["fusion"]["mode"],
                                   eps_temp=mm["fusion"].get("eps_temp", 0.1))
    out["multimodal"]["fusion_kwargs"] = dict(num_classes=mm["model"]["num_classes"], mode=mm["fusion"]["mode"],
                                              eps_temp=mm["fusion"].get("eps_temp", 0.1))
    out["multimodal"]["fusion"] = {"state_dict": [[k, list(v.shape)] for k, v in fz.state_dict().items()]}
    out["multimodal"]["training"] = {k: mm["training"][k] for k in ("encoder_learning_rate", "fusion_learning_rate",
                                     "weight_decay", "warmup_epochs", "epochs", "lambda_aux_img", "lambda_aux_eeg",
                                     "lambda_reg", "fp16", "max_grad_norm")}
    out["multimodal"]["freeze"] = {"gaze": mm["gaze_encoder"].get("freeze", False), "eeg": mm["eeg_encoder"].get("freeze", False)}
    path = os.path.join(GOLD, "constructor_contract.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
