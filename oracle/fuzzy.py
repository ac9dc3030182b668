"""Functional CPU restatement of FuzzyGatingFusion (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows 3_Models/fusion/fuzzy_gating_fusion.py (cited as ``fgf:<line>``).
"""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
MODES = ("full", "no_temperature", "no_fuzzification", "fixed_weights")  # fgf:61


def inverse_softplus(x: float) -> float:
    if x <= 0:
        raise ValueError("inverse_softplus requires x > 0")  # fgf:18-19
    return math.log(math.expm1(x))


def init_params(num_classes: int = 3, eps_temp: float = 0.1) -> Dict[str, Tensor]:
    """Initial parameter values of fgf:88-120 under the reference's state_dict key names."""
    hmax = math.log(num_classes)
    ls = math.log(hmax * 0.3)
    return {
        "tau_img": torch.tensor(inverse_softplus(1.5 - eps_temp)),
        "tau_eeg": torch.tensor(inverse_softplus(1.0 - eps_temp)),
        "c_reliable": torch.tensor(0.0),
        "c_unreliable_img": torch.tensor(hmax * 0.8),
        "c_unreliable_eeg": torch.tensor(hmax * 0.8),
        "log_sigma_reliable_img": torch.tensor(ls),
        "log_sigma_reliable_eeg": torch.tensor(ls),
        "log_sigma_unreliable_img": torch.tensor(ls),
        "log_sigma_unreliable_eeg": torch.tensor(ls),
        "beta": torch.tensor([math.log(0.8 / 0.2), math.log(0.2 / 0.8), math.log(0.6 / 0.4), 0.0]),
    }


def entropy(logits: Tensor, eps_log: float = 1e-8) -> Tensor:
    p = F.softmax(logits, dim=-1)
    return -(p * torch.log(p + eps_log)).sum(-1)  # fgf:142-145: log(p + eps), not log_softmax


def membership(h: Tensor, c: Tensor, log_sigma: Tensor, eps_div: float = 1e-8) -> Tensor:
    return torch.exp(-((h - c) ** 2) / (2 * torch.exp(log_sigma) ** 2 + eps_div))  # fgf:166-167


def fuzzy_forward(p: Dict[str, Tensor], img_logits: Tensor, eeg_logits: Tensor, mode: str = "full",
                  num_classes: int = 3, eps_temp: float = 0.1, eps_log: float = 1e-8,
                  eps_div: float = 1e-8) -> Tuple[Tensor, Tensor, Dict]:
    """FuzzyGatingFusion.forward (fgf:297-390) -> (fused_logits, alpha, aux_info)."""
    if mode not in MODES:
        raise ValueError(f"Invalid mode '{mode}'. Must be one of {MODES}")
    B = img_logits.shape[0]
    aux: Dict = {}
    if mode in ("no_temperature", "fixed_weights"):
        t_img, t_eeg = torch.ones(1), torch.ones(1)
        zi, ze = img_logits, eeg_logits
    else:
        t_img = F.softplus(p["tau_img"]) + eps_temp
        t_eeg = F.softplus(p["tau_eeg"]) + eps_temp
        zi, ze = img_logits / t_img, eeg_logits / t_eeg
    aux["temperatures"] = {"img": t_img.detach(), "eeg": t_eeg.detach()}
    hi, he = entropy(zi, eps_log), entropy(ze, eps_log)
    aux["entropies"] = {"img": hi.detach(), "eeg": he.detach()}
    aux["membership"] = aux["firing_strengths"] = aux["consequents"] = None
    if mode == "fixed_weights":
        alpha = torch.full((B,), 0.5)
    elif mode == "no_fuzzification":
        hmax = math.log(num_classes)
        ci = torch.clamp(1.0 - hi / (hmax + eps_div), min=0.0)
        ce = torch.clamp(1.0 - he / (hmax + eps_div), min=0.0)
        alpha = torch.clamp(ci / (ci + ce + eps_div), 0.0, 1.0)  # fgf:286-295
    else:
        mir = membership(hi, p["c_reliable"], p["log_sigma_reliable_img"], eps_div)
        miu = membership(hi, p["c_unreliable_img"], p["log_sigma_unreliable_img"], eps_div)
        mer = membership(he, p["c_reliable"], p["log_sigma_reliable_eeg"], eps_div)
        meu = membership(he, p["c_unreliable_eeg"], p["log_sigma_unreliable_eeg"], eps_div)
        w = torch.stack([mir * meu, miu * mer, mir * mer, miu * meu], dim=-1)  # fgf:231-236
        theta = torch.sigmoid(p["beta"])
        alpha = torch.clamp((w * theta).sum(-1) / (w.sum(-1) + eps_div), 0.0, 1.0)  # fgf:259-264
        aux["membership"] = {"img": {"rel": mir.detach(), "unrel": miu.detach()},
                             "eeg": {"rel": mer.detach(), "unrel": meu.detach()}}
        aux["firing_strengths"] = w.detach()
        aux["consequents"] = theta.detach()
    aux["fuzz_params"] = {
        "c_unreliable": {"img": p["c_unreliable_img"].detach(), "eeg": p["c_unreliable_eeg"].detach()},
        "sigma_reliable": {"img": torch.exp(p["log_sigma_reliable_img"]).detach(),
                           "eeg": torch.exp(p["log_sigma_reliable_eeg"]).detach()},
        "sigma_unreliable": {"img": torch.exp(p["log_sigma_unreliable_img"]).detach(),
                             "eeg": torch.exp(p["log_sigma_unreliable_eeg"]).detach()},
    }
    a = alpha.unsqueeze(-1)
    return a * zi + (1 - a) * ze, alpha, aux  # fgf:387-388


def temperature_regularization(p: Dict[str, Tensor], t_min: float = 0.5, t_max: float = 5.0,
                               eps_temp: float = 0.1) -> Tensor:
    """fgf:412-419."""
    ti = F.softplus(p["tau_img"]) + eps_temp
    te = F.softplus(p["tau_eeg"]) + eps_temp
    return F.relu(ti - t_max) + F.relu(t_min - ti) + F.relu(te - t_max) + F.relu(t_min - te)


def multimodal_loss(fused: Tensor, img_logits: Tensor, eeg_logits: Tensor, aux: Dict, reg: Tensor, labels: Tensor,
                    lambda_img: float = 0.3, lambda_eeg: float = 0.3, lambda_reg: float = 0.1) -> Tensor:
    """Loss composition of 4_Experiments/scripts/train_multimodal_fuzzy_fusion.py:440-460."""
    t = aux["temperatures"]
    return (F.cross_entropy(fused, labels) + lambda_img * F.cross_entropy(img_logits / t["img"], labels)
            + lambda_eeg * F.cross_entropy(eeg_logits / t["eeg"], labels) + lambda_reg * reg.squeeze())
