"""Aux-loss goldens (TEST INFRASTRUCTURE): the three optional batch-level losses of the UNMODIFIED reference
(3_Models/backbones/dual_eeg_transformer.py:1255-1371) on seeded (B, d) tokens, values and input gradients, incl. the
"no positive pairs" early return.  Run in the build container only:  python -m oracle.make_golden_aux"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = load_reference()
    m = ref.det.DualEEGTransformer(in_channels=4, d_model=16, num_layers=1, num_heads=2, d_ff=32, max_len=64,
                                   use_spectrogram=False, use_ibs=False)
    g = torch.Generator().manual_seed(99)
    B, D = 12, 48
    arrs = {}
    ibs = torch.randn(B, D, generator=g).requires_grad_(True)
    c1 = (torch.randn(B, D, generator=g) * 2).requires_grad_(True)
    c2 = (0.5 * c1.detach() + torch.randn(B, D, generator=g)).requires_grad_(True)
    labels = torch.tensor([0, 1, 2, 0, 1, 2, 2, 2, 0, 1, 1, 0])
    arrs.update(ibs=ibs.detach().numpy(), cls1=c1.detach().numpy(), cls2=c2.detach().numpy(), labels=labels.numpy())
    for name, fn in (("sym", lambda: m.compute_symmetry_loss(c1, c2)),
                     ("align", lambda: m.compute_ibs_alignment_loss(ibs, c1, c2)),
                     ("align_t05", lambda: m.compute_ibs_alignment_loss(ibs, c1, c2, temperature=0.5)),
                     ("contrast", lambda: m.compute_ibs_contrastive_loss(ibs, labels)),
                     ("contrast_t05", lambda: m.compute_ibs_contrastive_loss(ibs, labels, temperature=0.5))):
        for t in (ibs, c1, c2):
            t.grad = None
        loss = fn()
        loss.backward()
        arrs[name + "::loss"] = loss.detach().numpy()
        for tn, t in (("ibs", ibs), ("cls1", c1), ("cls2", c2)):
            if t.grad is not None:
                arrs[f"{name}::grad_{tn}"] = t.grad.numpy().copy()
    # a label that occurs once has no positive pair: that row is excluded from the mean (det:1366-1369)
    lab1 = torch.tensor([0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2])
    ibs.grad = None
    loss = m.compute_ibs_contrastive_loss(ibs, lab1)
    loss.backward()
    arrs.update({"contrast_single::labels": lab1.numpy(), "contrast_single::loss": loss.detach().numpy(),
                 "contrast_single::grad_ibs": ibs.grad.numpy().copy()})
    # no positives at all -> 0 (det:1349-1351)
    arrs["contrast_nopos::loss"] = m.compute_ibs_contrastive_loss(ibs.detach()[:3], torch.tensor([0, 1, 2])).numpy()
    path = os.path.join(GOLD, "aux_losses.npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: float(v) for k, v in arrs.items() if k.endswith("::loss")})


if __name__ == "__main__":
    main()
