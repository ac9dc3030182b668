"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
Each fixture stores the seeded inputs, the reference ``state_dict`` (keys prefixed ``sd::``) and the
reference outputs / gradients, so the oracle restatement and the CUDA path can be checked anywhere.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402
from eyegaze_multimodal_b200.synth import eeg_pair_batch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrs):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def golden_ibs(ref):
    for tag, (B, C, T) in {"small": (3, 8, 256), "c32": (1, 32, 1024)}.items():
        e1, e2 = eeg_pair_batch(B, C, T, seed=3, coupled=True)
        gen = ref.det.IBSConnectivityMatrixGenerator(C, 256, feature_type="all")
        with torch.no_grad():
            m = gen(e1, e2)
        _save(f"ibs_{tag}.npz", eeg1=_np(e1), eeg2=_np(e2), matrices=_np(m))


MODEL_CASES = {
    # name: (ctor kwargs, (B, C, T))
    "full": (dict(in_channels=8, num_classes=3, d_model=32, num_layers=2, num_heads=4, d_ff=64, dropout=0.1,
                  max_len=96), (4, 8, 256)),
    "a1_baseline": (dict(in_channels=8, num_classes=3, d_model=16, num_layers=1, num_heads=2, d_ff=32, max_len=64,
                         use_spectrogram=False, use_ibs=False), (3, 8, 256)),
    "scalar_ibs": (dict(in_channels=6, num_classes=3, d_model=16, num_layers=1, num_heads=2, d_ff=32, max_len=64,
                        use_spectrogram=False, use_robust_ibs=False), (3, 6, 256)),
    "phase_noin_nocross": (dict(in_channels=6, num_classes=3, d_model=16, num_layers=1, num_heads=4, d_ff=32,
                                max_len=64, ibs_feature_type="phase", ibs_instance_norm=False,
                                use_cross_attention=False), (3, 6, 256)),
}
GRAD_KEYS = ["temporal_conv.convs.0.weight", "encoder.layers.0.mha.q_proj.weight", "encoder.layers.0.ln1.weight",
             "classifier.3.bias", "cls_token", "pos_embed.pos_embed.weight"]


def golden_models(ref):
    for name, (kw, (B, C, T)) in MODEL_CASES.items():
        torch.manual_seed(1234)
        model = ref.det.DualEEGTransformer(**kw)
        model.eval()  # dropout off; autograd stays on (hard part 4)
        e1, e2 = eeg_pair_batch(B, C, T, seed=11, coupled=True)
        labels = torch.tensor([i % 3 for i in range(B)])
        out = model(e1, e2, labels)
        loss = out["loss"] + (out["loss_ibs_cls"] if "loss_ibs_cls" in out else 0.0)
        loss.backward()
        arrs = {"eeg1": _np(e1), "eeg2": _np(e2), "labels": _np(labels)}
        for k, v in model.state_dict().items():
            arrs["sd::" + k] = _np(v)
        for k, v in out.items():
            if v is not None:
                arrs["out::" + k] = _np(v)
        params = dict(model.named_parameters())
        for k in GRAD_KEYS:
            if k in params and params[k].grad is not None:
                arrs["grad::" + k] = _np(params[k].grad)
        extra = [k for k in ("ibs_tokenizer.bottleneck.0.weight", "spectrogram_generator.spec_conv.0.weight",
                             "spectrogram_generator.spec_conv.3.weight", "ibs_tokenizer.instance_norm.weight",
                             "cross_attn.cross_attn.v_proj.weight", "ibs_generator.proj.0.weight") if k in params]
        for k in extra:
            if params[k].grad is not None:
                arrs["grad::" + k] = _np(params[k].grad)
        if hasattr(model, "ibs_matrix_generator"):
            with torch.no_grad():
                arrs["ibs_matrices"] = _np(model.ibs_matrix_generator(e1, e2))
        arrs["kwargs_repr"] = np.array(repr(kw))
        _save(f"eeg_model_{name}.npz", **arrs)


def golden_mha(ref):
    torch.manual_seed(7)
    m = ref.det.CrossBrainAttention(64, 4, dropout=0.1).eval()
    z1 = torch.randn(3, 16, 64, requires_grad=True)
    z2 = torch.randn(3, 24, 64, requires_grad=True)
    mha = m.cross_attn
    o = mha(z1, z2, z2)   # Lq != Lk is legal (art.py:203-205), BASELINE config 3
    (o ** 2).sum().backward()
    arrs = {"z1": _np(z1), "z2": _np(z2), "out": _np(o), "grad_z1": _np(z1.grad), "grad_z2": _np(z2.grad)}
    for k, v in m.state_dict().items():
        arrs["sd::" + k] = _np(v)
    for k, p in m.named_parameters():
        if p.grad is not None:
            arrs["grad::" + k] = _np(p.grad)
    a = torch.randn(3, 20, 64)
    b = torch.randn(3, 20, 64)
    with torch.no_grad():
        o1, o2 = m(a, b)
    arrs.update(a=_np(a), b=_np(b), cross1=_np(o1), cross2=_np(o2))
    _save("cross_attention.npz", **arrs)


def golden_fuzzy(ref):
    torch.manual_seed(5)
    img = torch.randn(8, 3) * 2
    eeg = torch.randn(8, 3) * 2
    arrs = {"img": _np(img), "eeg": _np(eeg)}
    for mode in ("full", "no_temperature", "no_fuzzification", "fixed_weights"):
        m = ref.fgf.FuzzyGatingFusion(3, mode=mode)
        i = img.clone().requires_grad_(True)
        e = eeg.clone().requires_grad_(True)
        fused, alpha, aux = m(i, e)
        (fused * torch.arange(1, 4)).sum().backward()
        arrs[f"{mode}::fused"] = _np(fused)
        arrs[f"{mode}::alpha"] = _np(alpha)
        arrs[f"{mode}::H_img"] = _np(aux["entropies"]["img"])
        arrs[f"{mode}::H_eeg"] = _np(aux["entropies"]["eeg"])
        arrs[f"{mode}::grad_img"] = _np(i.grad)
        arrs[f"{mode}::grad_eeg"] = _np(e.grad)
        if aux["firing_strengths"] is not None:
            arrs[f"{mode}::firing"] = _np(aux["firing_strengths"])
        for k, p in m.named_parameters():
            arrs[f"{mode}::grad::{k}"] = _np(p.grad) if p.grad is not None else np.zeros(p.shape, np.float32)
    m = ref.fgf.FuzzyGatingFusion(3, mode="full")
    uni = torch.zeros(8, 3)
    conf = torch.tensor([[10.0, -10.0, -10.0]] * 8)
    with torch.no_grad():   # the reference self-test's deterministic edge cases (fgf:498-509)
        arrs["edge::uniform"] = _np(m(uni, uni)[1])
        arrs["edge::conf_img"] = _np(m(conf, uni)[1])
        arrs["edge::conf_eeg"] = _np(m(uni, conf)[1])
        arrs["edge::t_img"] = _np(m.temp_img)
        arrs["edge::t_eeg"] = _np(m.temp_eeg)
        arrs["edge::reg"] = _np(m.compute_temperature_regularization())
    _save("fuzzy_fusion.npz", **arrs)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference()
    golden_fuzzy(ref)
    golden_mha(ref)
    golden_ibs(ref)
    golden_models(ref)


if __name__ == "__main__":
    main()
