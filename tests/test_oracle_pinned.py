"""Pins the CPU oracle (oracle/) to golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py) and to the reference self-test's known answers. CPU only."""
import ast
import math

import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden
from oracle import eeg as O
from oracle import fuzzy as FZ


def ibs_flip_aware_compare(got, ref, T, smooth_tol=2e-5, max_flips_frac=0.02):
    """PLI (1) / wPLI (2) may differ by a few sign() flips; the other five
    features must agree to smooth_tol."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    for f in (0, 3, 4, 5, 6):
        d = np.abs(got[:, :, f] - ref[:, :, f]).max()
        assert d <= smooth_tol, f"feature {f}: max abs err {d}"
    dpli = np.abs(got[:, :, 1] - ref[:, :, 1])
    flipped = dpli > 1e-5
    assert flipped.mean() <= max_flips_frac, f"too many PLI entries differ: {flipped.mean()}"
    assert (dpli <= 8.0 / T + 1e-6).all(), "PLI differs by more than 4 sign flips"
    dw = np.abs(got[:, :, 2] - ref[:, :, 2])
    assert (dw > 1e-4).mean() <= max_flips_frac and dw.max() < 0.05


@pytest.mark.parametrize("name", ["ibs_small.npz", "ibs_c32.npz"])
def test_ibs_connectivity_matches_reference(name):
    g = load_golden(name)
    e1, e2 = torch.from_numpy(g["eeg1"]), torch.from_numpy(g["eeg2"])
    got = O.ibs_connectivity(e1, e2, 256.0, "all")
    ibs_flip_aware_compare(got.numpy(), g["matrices"], e1.shape[-1])


def test_ibs_coherence_is_band_fraction():
    """Known answer from SURVEY.md 8a-3: single-segment coherence = (#in-band bins)/(T/2+1)."""
    g = load_golden("ibs_c32.npz")
    coh = g["matrices"][:, :, 3]
    for b, frac in enumerate([179, 15, 17, 21, 69, 61]):
        assert np.allclose(coh[:, b], frac / 513.0, atol=2e-6)


def _cfg_from(g):
    kw = ast.literal_eval(str(g["kwargs_repr"]))
    return O.EEGConfig(**kw)


@pytest.mark.parametrize("name", ["full", "a1_baseline", "scalar_ibs", "phase_noin_nocross"])
def test_dual_eeg_forward_and_grads_match_reference(name):
    g = load_golden(f"eeg_model_{name}.npz")
    cfg = _cfg_from(g)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in golden_state_dict(g).items()}
    e1, e2 = torch.from_numpy(g["eeg1"]), torch.from_numpy(g["eeg2"])
    labels = torch.from_numpy(g["labels"])
    mats = torch.from_numpy(g["ibs_matrices"]) if "ibs_matrices" in g else None
    out = O.dual_eeg_forward(sd, e1, e2, cfg, labels, ibs_matrices=mats)
    for k in ("logits", "cls1", "cls2", "ibs_logits", "ibs_token", "loss", "loss_ibs_cls"):
        if "out::" + k in g:
            np.testing.assert_allclose(out[k].detach().numpy(), g["out::" + k], atol=2e-6, rtol=1e-5, err_msg=k)
    loss = out["loss"] + (out["loss_ibs_cls"] if "loss_ibs_cls" in out else 0.0)
    loss.backward()
    n = 0
    for k, v in g.items():
        if k.startswith("grad::"):
            gr = sd[k[6:]].grad
            assert gr is not None, k
            np.testing.assert_allclose(gr.numpy(), v, atol=3e-6, rtol=2e-4, err_msg=k)
            n += 1
    assert n >= 5
    # own-generator path (no cached matrices) stays within the logits gate
    if mats is not None:
        out2 = O.dual_eeg_forward({k: v.detach() for k, v in sd.items()}, e1, e2, cfg)
        assert (out2["logits"] - torch.from_numpy(g["out::logits"])).abs().max() < 1e-5


def test_cross_attention_unequal_lengths():
    g = load_golden("cross_attention.npz")
    sd = {k: v.clone().requires_grad_(True) for k, v in golden_state_dict(g).items()}
    z1 = torch.from_numpy(g["z1"]).requires_grad_(True)
    z2 = torch.from_numpy(g["z2"]).requires_grad_(True)
    o = O.mha(z1, z2, z2, sd, "cross_attn.", 4)
    np.testing.assert_allclose(o.detach().numpy(), g["out"], atol=2e-6)
    (o ** 2).sum().backward()
    np.testing.assert_allclose(z1.grad.numpy(), g["grad_z1"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(z2.grad.numpy(), g["grad_z2"], atol=1e-5, rtol=1e-4)
    for k, v in g.items():
        if k.startswith("grad::"):
            np.testing.assert_allclose(sd[k[6:]].grad.numpy(), v, atol=2e-5, rtol=1e-4, err_msg=k)
    cfg = O.EEGConfig(d_model=64, num_heads=4)
    sdd = {k: v.detach() for k, v in sd.items()}
    o1, o2 = O.cross_brain(torch.from_numpy(g["a"]), torch.from_numpy(g["b"]), sdd, cfg, pre="")
    np.testing.assert_allclose(o1.numpy(), g["cross1"], atol=3e-6)
    np.testing.assert_allclose(o2.numpy(), g["cross2"], atol=3e-6)


@pytest.mark.parametrize("mode", ["full", "no_temperature", "no_fuzzification", "fixed_weights"])
def test_fuzzy_fusion_matches_reference(mode):
    g = load_golden("fuzzy_fusion.npz")
    p = {k: v.clone().requires_grad_(k != "c_reliable") for k, v in FZ.init_params().items()}
    img = torch.from_numpy(g["img"]).requires_grad_(True)
    eeg = torch.from_numpy(g["eeg"]).requires_grad_(True)
    fused, alpha, aux = FZ.fuzzy_forward(p, img, eeg, mode)
    np.testing.assert_allclose(fused.detach().numpy(), g[f"{mode}::fused"], atol=1e-6)
    np.testing.assert_allclose(alpha.detach().numpy(), g[f"{mode}::alpha"], atol=1e-6)
    np.testing.assert_allclose(aux["entropies"]["img"].numpy(), g[f"{mode}::H_img"], atol=1e-6)
    (fused * torch.arange(1, 4)).sum().backward()
    np.testing.assert_allclose(img.grad.numpy(), g[f"{mode}::grad_img"], atol=2e-6)
    np.testing.assert_allclose(eeg.grad.numpy(), g[f"{mode}::grad_eeg"], atol=2e-6)
    for k in p:
        key = f"{mode}::grad::{k}"
        if key in g:
            got = p[k].grad.numpy() if p[k].grad is not None else np.zeros_like(g[key])
            np.testing.assert_allclose(got, g[key], atol=2e-6, err_msg=k)
    assert set(aux) == {"temperatures", "entropies", "membership", "firing_strengths", "consequents", "fuzz_params"}


def test_fuzzy_known_answers_from_reference_selftest():
    """fuzzy_gating_fusion.py:489-517 + SURVEY.md 4: alpha 0.5 / 0.7907 / 0.2102, T 1.5 / 1.0, reg 0."""
    g = load_golden("fuzzy_fusion.npz")
    p = FZ.init_params()
    uni = torch.zeros(8, 3)
    conf = torch.tensor([[10.0, -10.0, -10.0]] * 8)
    a0 = FZ.fuzzy_forward(p, uni, uni)[1]
    a1 = FZ.fuzzy_forward(p, conf, uni)[1]
    a2 = FZ.fuzzy_forward(p, uni, conf)[1]
    assert torch.allclose(a0, torch.full((8,), 0.5), atol=1e-6)
    assert abs(a1.mean().item() - 0.7907) < 5e-5 and abs(a2.mean().item() - 0.2102) < 5e-5
    np.testing.assert_allclose(a1.numpy(), g["edge::conf_img"], atol=1e-6)
    np.testing.assert_allclose(a2.numpy(), g["edge::conf_eeg"], atol=1e-6)
    assert torch.allclose(FZ.fuzzy_forward(p, conf, uni, "fixed_weights")[1], torch.full((8,), 0.5))
    t = FZ.fuzzy_forward(p, uni, uni)[2]["temperatures"]
    assert abs(t["img"].item() - 1.5) < 1e-6 and abs(t["eeg"].item() - 1.0) < 1e-6
    assert FZ.temperature_regularization(p).item() == 0.0
    with pytest.raises(ValueError):
        FZ.fuzzy_forward(p, uni, uni, "bogus")
    with pytest.raises(ValueError):
        FZ.inverse_softplus(0.0)


# ---------------------------------------------------------------------------------------------------------------------
# Non-degenerate fixture: a state_dict TRAINED by the unmodified reference on the planted task, whose
# evaluation predictions cover all three classes with some errors; logits / argmax / ClassificationMetrics were written by
# the unmodified reference (oracle/make_golden_trained.py).
# ---------------------------------------------------------------------------------------------------------------------
def test_trained_fixture_is_not_degenerate():
    g = load_golden("eeg_model_trained.npz")
    assert set(g["preds"].tolist()) == {0, 1, 2}
    acc = float((g["preds"] == g["labels"]).mean())
    assert 0.5 < acc < 0.97
    top2 = np.sort(g["out::logits"], axis=-1)
    assert (top2[:, -1] - top2[:, -2]).min() >= 0.05          # argmax is robust to 1e-4 (fp32) and 2e-2 (bf16) errors


def test_oracle_matches_trained_reference_model_and_metrics():
    from oracle import metrics as M
    g = load_golden("eeg_model_trained.npz")
    cfg = _cfg_from(g)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in golden_state_dict(g).items()}
    e1, e2 = torch.from_numpy(g["eeg1"]), torch.from_numpy(g["eeg2"])
    labels = torch.from_numpy(g["labels"])
    out = O.dual_eeg_forward(sd, e1, e2, cfg, labels)
    assert np.abs(out["logits"].detach().numpy() - g["out::logits"]).max() <= 1e-4
    assert np.abs(out["ibs_logits"].detach().numpy() - g["out::ibs_logits"]).max() <= 1e-4
    preds = out["logits"].argmax(-1).numpy()
    assert (preds == g["preds"]).all()
    # oracle/metrics.py == the reference's ClassificationMetrics on the same predictions (exact: both are sklearn calls)
    want = dict(zip([str(k) for k in g["metric_names"]], g["metric_values"]))
    got = M.compute_metrics(g["labels"], preds)
    assert set(got) == set(want)
    for k, v in want.items():
        assert got[k] == pytest.approx(float(v), abs=1e-12), k
    assert (M.compute_confusion_matrix(g["labels"], preds) == g["confusion"]).all()
    probs = torch.softmax(out["logits"].detach(), -1).numpy()
    aucs = M.compute_aucs(g["labels"], probs)
    for k, v in zip([str(k) for k in g["auc_names"]], g["auc_values"]):
        assert aucs[k] == pytest.approx(float(v), abs=1e-6), k
    (out["loss"] + out["loss_ibs_cls"]).backward()
    for k, v in g.items():
        if k.startswith("grad::"):
            e = np.abs(sd[k[6:]].grad.numpy() - v).max()
            assert e <= 2e-5 + 2e-3 * np.abs(v).max(), (k, e)


def test_aux_losses_match_reference():
    """oracle/eeg.py aux losses (det:1255-1371) against values + gradients from the unmodified reference."""
    g = load_golden("aux_losses.npz")
    labels = torch.from_numpy(g["labels"])

    def leafs():
        return [torch.from_numpy(g[k]).clone().requires_grad_(True) for k in ("ibs", "cls1", "cls2")]
    cases = {"sym": lambda i, a, b: O.symmetry_loss(a, b), "align": lambda i, a, b: O.ibs_alignment_loss(i, a, b),
             "align_t05": lambda i, a, b: O.ibs_alignment_loss(i, a, b, 0.5),
             "contrast": lambda i, a, b: O.ibs_contrastive_loss(i, labels),
             "contrast_t05": lambda i, a, b: O.ibs_contrastive_loss(i, labels, 0.5),
             "contrast_single": lambda i, a, b: O.ibs_contrastive_loss(i, torch.from_numpy(g["contrast_single::labels"]))}
    for name, fn in cases.items():
        i, a, b = leafs()
        loss = fn(i, a, b)
        loss.backward()
        assert abs(float(loss) - float(g[name + "::loss"])) <= 1e-6 * max(1.0, abs(float(g[name + "::loss"]))), name
        for tn, t in (("ibs", i), ("cls1", a), ("cls2", b)):
            if f"{name}::grad_{tn}" in g:
                assert np.abs(t.grad.numpy() - g[f"{name}::grad_{tn}"]).max() <= 1e-6, (name, tn)
    assert float(O.ibs_contrastive_loss(torch.from_numpy(g["ibs"])[:3], torch.tensor([0, 1, 2]))) == float(g["contrast_nopos::loss"]) == 0.0
