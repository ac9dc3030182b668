"""Non-degenerate parity fixture (TEST INFRASTRUCTURE): trains the UNMODIFIED reference
DualEEGTransformer on CPU for a few hundred steps on a planted synthetic task (class-dependent coupling between the
two players, eyegaze_multimodal_b200.synth.eeg_pair_batch(coupled=True)) so that the evaluation predictions cover
all three classes, then records -- all from the unmodified reference code --

  * the trained ``state_dict``, the evaluation inputs / labels,
  * reference logits, probabilities, argmax predictions,
  * ``ClassificationMetrics.compute_metrics`` / ``compute_confusion_matrix`` / per-class ROC-AUC of
    /root/reference/5_Metrics/classification_metrics.py on those predictions.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden_trained
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import REFERENCE_ROOT, load_reference  # noqa: E402
from eyegaze_multimodal_b200.synth import _zscore, eeg_pair_batch, gen_eeg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
KW = dict(in_channels=8, num_classes=3, d_model=32, num_layers=2, num_heads=4, d_ff=64, dropout=0.1, max_len=96)
B_TRAIN, B_EVAL, T, STEPS = 12, 36, 256, 150
MIN_MARGIN = 0.05      # evaluation trials whose reference top-2 margin is smaller are dropped (argmax must be robust)


def load_metrics_module():
    path = os.path.join(REFERENCE_ROOT, "5_Metrics", "classification_metrics.py")
    spec = importlib.util.spec_from_file_location("ref_classification_metrics", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def eval_batch(B, C, T, seed):
    """Evaluation trials with the SAME planted structure as eeg_pair_batch(coupled=True) but a coupling strength that
    fades from 0.4 (the training value) to 0.03, so that the trained reference model gets some of them wrong and the
    metrics are not all 1.0."""
    e1 = np.empty((B, C, T), dtype=np.float32)
    e2 = np.empty((B, C, T), dtype=np.float32)
    for i in range(B):
        a = gen_eeg(C, T, 256.0, seed=seed * 100003 + i)
        b = gen_eeg(C, T, 256.0, seed=seed * 100019 + i)
        c = 0.4 - 0.37 * (i // 3) / max(B // 3 - 1, 1)
        k = i % 3
        if k == 1:
            b = (1 - c) * b + c * a
        elif k == 2:
            b = (1 - c) * b + c * np.roll(a, 16, axis=1)
        e1[i], e2[i] = _zscore(a), _zscore(b)
    return torch.from_numpy(e1), torch.from_numpy(e2)


def main():
    ref = load_reference()
    cm = load_metrics_module()
    torch.manual_seed(4321)
    model = ref.det.DualEEGTransformer(**KW)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=0.01)   # train_art.py:388-398
    model.train()
    for step in range(STEPS):
        e1, e2 = eeg_pair_batch(B_TRAIN, KW["in_channels"], T, seed=1000 + step, coupled=True)
        labels = torch.arange(B_TRAIN) % 3                     # the label IS the planted coupling class
        out = model(e1, e2, labels)
        loss = out["loss"] + 0.5 * out["loss_ibs_cls"]
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)   # train_art.py:221
        opt.step()
        if step % 20 == 0:
            print("step %d loss %.4f" % (step, float(loss)), flush=True)
    model.eval()
    e1, e2 = eval_batch(B_EVAL, KW["in_channels"], T, seed=77)
    labels = torch.arange(B_EVAL) % 3
    with torch.no_grad():
        lg = model(e1, e2)["logits"]
    top2 = torch.sort(lg, dim=-1).values
    keep = (top2[:, -1] - top2[:, -2]) >= MIN_MARGIN
    print("kept %d of %d evaluation trials (margin >= %.2f)" % (int(keep.sum()), B_EVAL, MIN_MARGIN))
    e1, e2, labels = e1[keep].contiguous(), e2[keep].contiguous(), labels[keep].contiguous()
    model.zero_grad(set_to_none=True)                      # drop the last training step's gradients
    out = model(e1, e2, labels)
    (out["loss"] + out["loss_ibs_cls"]).backward()
    logits = out["logits"].detach()
    probs = torch.softmax(logits, dim=-1).numpy()
    preds = logits.argmax(-1).numpy()
    y = labels.numpy()
    print("predictions:", preds, "labels:", y)
    assert set(preds.tolist()) == {0, 1, 2}, "fixture is degenerate: train longer"
    acc = float((preds == y).mean())
    assert 0.5 < acc < 0.97, "fixture should be non-trivial (accuracy %.3f)" % acc
    margins = np.sort(logits.numpy(), axis=-1)
    print("min top-2 margin: %.4f" % float((margins[:, -1] - margins[:, -2]).min()))
    calc = cm.ClassificationMetrics()
    metrics = calc.compute_metrics(y, preds)
    conf = calc.compute_confusion_matrix(y, preds)
    roc = calc.compute_roc_data(y, probs)
    print({k: round(float(v), 4) for k, v in metrics.items()})
    arrs = {"eeg1": e1.numpy(), "eeg2": e2.numpy(), "labels": y, "kwargs_repr": np.array(repr(KW)),
            "out::logits": logits.numpy(), "out::ibs_logits": out["ibs_logits"].detach().numpy(),
            "out::loss": out["loss"].detach().numpy(), "out::loss_ibs_cls": out["loss_ibs_cls"].detach().numpy(),
            "probs": probs, "preds": preds,
            "metric_names": np.array(sorted(metrics)), "metric_values": np.array([metrics[k] for k in sorted(metrics)]),
            "confusion": conf,
            "auc_names": np.array(list(calc.class_names) + ["micro", "macro"]),
            "auc_values": np.array([roc[k]["auc"] for k in list(calc.class_names) + ["micro", "macro"]])}
    for k, v in model.state_dict().items():
        arrs["sd::" + k] = v.detach().numpy()
    params = dict(model.named_parameters())
    for k in ("temporal_conv.convs.0.weight", "encoder.layers.1.ffn.linear2.weight", "classifier.3.weight",
              "cls_token", "pos_embed.pos_embed.weight", "ibs_tokenizer.bottleneck.0.weight"):
        arrs["grad::" + k] = params[k].grad.detach().numpy()
    path = os.path.join(GOLD, "eeg_model_trained.npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
