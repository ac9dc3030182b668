"""Synthetic inputs of the shapes BASELINE.json names (no dataset ships with the reference).

EEG follows the reference's own generator recipe ``gen_eeg`` (1_Data/processed/two_EEG_fusion.py:31-49,
mode='mixed': per-channel sum of 3 sinusoids 1-40 Hz, amplitude U(0.1,1), phase U(0,2pi), plus N(0,0.1^2)
noise; seeds ``s*100003+i`` / ``s*100019+i`` as at :62-63) followed by the dataset's global z-score
``(x - x.mean()) / (x.std() + 1e-8)`` with NumPy's population std (1_Data/processed/dual_eeg_dataset.py:201-202).
Gaze heat-maps are ``randn`` (ImageNet-normalised images are ~zero-mean/unit-variance; the reference's own
smoke test uses the same, early_fusion_vit.py:288-289); labels are ``randint(0, 3)``.
"""
import numpy as np
import torch


def gen_eeg(C: int = 32, T: int = 1024, sample_rate: float = 256.0, noise_std: float = 0.1,
            num_components: int = 3, seed=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(T, dtype=np.float32) / float(sample_rate)
    x = np.zeros((C, T), dtype=np.float32)
    for c in range(C):
        freqs = rng.uniform(1.0, 40.0, size=(num_components,)).astype(np.float32)
        amps = rng.uniform(0.1, 1.0, size=(num_components,)).astype(np.float32)
        phases = rng.uniform(0.0, 2.0 * np.pi, size=(num_components,)).astype(np.float32)
        s = np.zeros_like(t)
        for f, a, p in zip(freqs, amps, phases):
            s += a * np.sin(2.0 * np.pi * f * t + p)
        x[c] += s.astype(np.float32)
    x += rng.normal(loc=0.0, scale=noise_std, size=(C, T)).astype(np.float32)
    return x


def _zscore(x: np.ndarray) -> np.ndarray:
    return ((x - x.mean()) / (x.std() + 1e-8)).astype(np.float32)


def eeg_pair_batch(B: int, C: int, T: int, seed: int = 0, sample_rate: float = 256.0, coupled: bool = False):
    """(eeg1, eeg2) float32 CPU tensors of shape (B, C, T).

    ``coupled=True`` plants class-dependent phase coupling between the two players (used by the
    non-degenerate argmax fixture: random-init models otherwise predict a single class).
    """
    e1 = np.empty((B, C, T), dtype=np.float32)
    e2 = np.empty((B, C, T), dtype=np.float32)
    for i in range(B):
        a = gen_eeg(C, T, sample_rate, seed=seed * 100003 + i)
        b = gen_eeg(C, T, sample_rate, seed=seed * 100019 + i)
        if coupled:
            k = i % 3
            if k == 1:
                b = 0.6 * b + 0.4 * a
            elif k == 2:
                b = 0.6 * b + 0.4 * np.roll(a, 16, axis=1)
        e1[i], e2[i] = _zscore(a), _zscore(b)
    return torch.from_numpy(e1), torch.from_numpy(e2)


def randn_eeg_pair(B: int, C: int, T: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, T, generator=g), torch.randn(B, C, T, generator=g)


def gaze_pair_batch(B: int, img: int = 224, seed: int = 0):
    g = torch.Generator().manual_seed(seed + 7919)
    return torch.randn(B, 3, img, img, generator=g), torch.randn(B, 3, img, img, generator=g)


def labels_batch(B: int, num_classes: int = 3, seed: int = 0):
    g = torch.Generator().manual_seed(seed + 104729)
    return torch.randint(0, num_classes, (B,), generator=g)
