"""CPU restatement of the reference's classification metrics (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows /root/reference/5_Metrics/classification_metrics.py: ``ClassificationMetrics.compute_metrics`` (:66-131),
``compute_confusion_matrix`` (:133-152) and the per-class / micro / macro AUCs of ``compute_roc_data`` (:154-230).
The reference delegates the arithmetic to scikit-learn (installed here and on the GPU box); what is restated is WHICH
sklearn calls are made with which arguments.  Pinned by tests/golden/eeg_model_trained.npz, whose metric values were
written by the unmodified reference class (oracle/make_golden_trained.py).
"""
from typing import Dict, List

import numpy as np
from sklearn.metrics import (accuracy_score, auc, confusion_matrix, f1_score, precision_score, recall_score, roc_curve)
from sklearn.preprocessing import label_binarize

DEFAULT_CLASS_NAMES = ["Single", "Competition", "Cooperation"]          # classification_metrics.py:29


def compute_metrics(y_true: np.ndarray, y_pred: np.ndarray, class_names: List[str] = None) -> Dict[str, float]:
    names = class_names or DEFAULT_CLASS_NAMES
    m = {"accuracy": accuracy_score(y_true, y_pred)}
    for avg in ("macro", "weighted"):                                   # :95-114
        m[f"precision_{avg}"] = precision_score(y_true, y_pred, average=avg, zero_division=0)
        m[f"recall_{avg}"] = recall_score(y_true, y_pred, average=avg, zero_division=0)
        m[f"f1_{avg}"] = f1_score(y_true, y_pred, average=avg, zero_division=0)
    for i, name in enumerate(names):                                    # :117-129, one-vs-rest per class
        t, p = (y_true == i).astype(int), (y_pred == i).astype(int)
        m[f"precision_{name}"] = precision_score(t, p, zero_division=0)
        m[f"recall_{name}"] = recall_score(t, p, zero_division=0)
        m[f"f1_{name}"] = f1_score(t, p, zero_division=0)
    return m


def compute_confusion_matrix(y_true: np.ndarray, y_pred: np.ndarray, n_classes: int = 3) -> np.ndarray:
    return confusion_matrix(y_true, y_pred, labels=range(n_classes))   # :152


def compute_aucs(y_true: np.ndarray, y_prob: np.ndarray, class_names: List[str] = None) -> Dict[str, float]:
    """AUC entries of compute_roc_data: per class (one-vs-rest), 'micro' (ravelled), 'macro' (mean interpolated TPR)."""
    names = class_names or DEFAULT_CLASS_NAMES
    n = len(names)
    yb = label_binarize(y_true, classes=range(n))
    if n == 2:
        yb = np.hstack([1 - yb, yb])
    out, fprs, tprs = {}, [], []
    for i, name in enumerate(names):
        fpr, tpr, _ = roc_curve(yb[:, i], y_prob[:, i])
        out[name] = auc(fpr, tpr)
        fprs.append(fpr)
        tprs.append(tpr)
    fpr, tpr, _ = roc_curve(yb.ravel(), y_prob.ravel())
    out["micro"] = auc(fpr, tpr)
    all_fpr = np.unique(np.concatenate(fprs))
    mean_tpr = np.zeros_like(all_fpr)
    for f, t in zip(fprs, tprs):
        mean_tpr += np.interp(all_fpr, f, t)
    out["macro"] = auc(all_fpr, mean_tpr / n)
    return out
