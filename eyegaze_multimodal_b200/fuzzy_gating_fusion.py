"""FuzzyGatingFusion -- drop-in for ``3_Models/fusion/fuzzy_gating_fusion.py`` (cited ``fgf:<line>``).

Same constructor, parameter / buffer names (``state_dict`` keys), ``forward`` return triple and ``aux_info`` keys;
the six stages (temperature, entropy, fuzzification, inference, defuzzification, fusion) and their backward run as
one kernel each instead of ~40 ATen micro-launches.
"""
import math
from typing import Dict, Literal, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def inverse_softplus(x: float) -> float:
    if x <= 0:
        raise ValueError("inverse_softplus requires x > 0")
    return math.log(math.expm1(x))


class FuzzyGatingFusion(nn.Module):
    VALID_MODES = ('full', 'no_temperature', 'no_fuzzification', 'fixed_weights')

    def __init__(self, num_classes: int = 3,
                 mode: Literal['full', 'no_temperature', 'no_fuzzification', 'fixed_weights'] = 'full',
                 eps_temp: float = 0.1, eps_log: float = 1e-8, eps_div: float = 1e-8):
        super().__init__()
        if mode not in self.VALID_MODES:
            raise ValueError(f"Invalid mode '{mode}'. Must be one of {self.VALID_MODES}")
        self.num_classes = num_classes
        self.mode = mode
        self.eps_temp = eps_temp
        self.eps_log = eps_log
        self.eps_div = eps_div
        self.max_entropy = math.log(num_classes)
        self.tau_img = nn.Parameter(torch.tensor(inverse_softplus(1.5 - eps_temp)))       # fgf:88-89
        self.tau_eeg = nn.Parameter(torch.tensor(inverse_softplus(1.0 - eps_temp)))
        self.register_buffer('c_reliable', torch.tensor(0.0))
        c_unrel_init = self.max_entropy * 0.8
        self.c_unreliable_img = nn.Parameter(torch.tensor(c_unrel_init))
        self.c_unreliable_eeg = nn.Parameter(torch.tensor(c_unrel_init))
        log_sigma_init = math.log(self.max_entropy * 0.3)
        self.log_sigma_reliable_img = nn.Parameter(torch.tensor(log_sigma_init))
        self.log_sigma_reliable_eeg = nn.Parameter(torch.tensor(log_sigma_init))
        self.log_sigma_unreliable_img = nn.Parameter(torch.tensor(log_sigma_init))
        self.log_sigma_unreliable_eeg = nn.Parameter(torch.tensor(log_sigma_init))
        self.beta = nn.Parameter(torch.tensor([math.log(0.8 / 0.2), math.log(0.2 / 0.8), math.log(0.6 / 0.4), 0.0]))

    @property
    def temp_img(self) -> torch.Tensor:
        return F.softplus(self.tau_img) + self.eps_temp

    @property
    def temp_eeg(self) -> torch.Tensor:
        return F.softplus(self.tau_eeg) + self.eps_temp

    def _params(self):
        return [self.tau_img, self.tau_eeg, self.c_reliable, self.c_unreliable_img, self.c_unreliable_eeg,
                self.log_sigma_reliable_img, self.log_sigma_reliable_eeg, self.log_sigma_unreliable_img,
                self.log_sigma_unreliable_eeg, self.beta]

    def forward(self, img_logits: torch.Tensor, eeg_logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, Dict]:
        B = img_logits.size(0)
        fused, alpha, aux = ops.fuzzy_gating(img_logits, eeg_logits, ops.FUZZY_MODES[self.mode], self.eps_temp,
                                             self.eps_log, self.eps_div, self._params())
        row, par = aux[:B], aux[B]                                  # detached analysis values written by the kernel
        no_temp = self.mode in ('no_temperature', 'fixed_weights')
        aux_info: Dict = {}
        aux_info['temperatures'] = {'img': par[8:9] if no_temp else par[8], 'eeg': par[9:10] if no_temp else par[9]}
        aux_info['entropies'] = {'img': row[:, 0], 'eeg': row[:, 1]}
        if self.mode in ('fixed_weights', 'no_fuzzification'):
            aux_info['membership'] = None
            aux_info['firing_strengths'] = None
            aux_info['consequents'] = None
        else:
            aux_info['membership'] = {'img': {'rel': row[:, 2], 'unrel': row[:, 3]},
                                      'eeg': {'rel': row[:, 4], 'unrel': row[:, 5]}}
            aux_info['firing_strengths'] = row[:, 6:10]
            aux_info['consequents'] = par[4:8]
        aux_info['fuzz_params'] = {
            'c_unreliable': {'img': self.c_unreliable_img.detach(), 'eeg': self.c_unreliable_eeg.detach()},
            'sigma_reliable': {'img': par[0], 'eeg': par[1]},
            'sigma_unreliable': {'img': par[2], 'eeg': par[3]},
        }
        return fused, alpha, aux_info

    def compute_temperature_regularization(self, t_min: float = 0.5, t_max: float = 5.0) -> torch.Tensor:
        """fgf:392-419: four scalar hinge terms on two 0-dim parameters."""
        T_img, T_eeg = self.temp_img, self.temp_eeg
        return F.relu(T_img - t_max) + F.relu(t_min - T_img) + F.relu(T_eeg - t_max) + F.relu(t_min - T_eeg)

    def extra_repr(self) -> str:
        return f"num_classes={self.num_classes}, mode='{self.mode}', eps_temp={self.eps_temp}"
