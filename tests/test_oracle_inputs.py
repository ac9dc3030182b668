"""CPU: the input-side oracle (oracle/inputs.py) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden_inputs.py: DualEEGDataset._preprocess_eeg, the enable_preprocessing=False expression, torchvision
ToTensor + Normalize as composed in multimodal_dataset.py:73-83)."""
import os

import numpy as np

from oracle import inputs as OI

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "inputs_golden.npz"))


def test_preprocess_eeg_matches_reference():
    got = OI.preprocess_eeg(G["eeg"])
    assert np.abs(got - G["eeg_preprocessed"]).max() <= 1e-6
    # properties the reference's docstring promises: zero mean / unit variance per channel, zero mean across channels
    assert np.abs(got.mean(axis=1)).max() <= 1e-5 and np.abs(got.std(axis=1) - 1).max() <= 1e-4


def test_simple_normalize_matches_reference():
    assert np.abs(OI.simple_normalize(G["eeg"]) - G["eeg_simple"]).max() <= 1e-6


def test_to_tensor_normalize_matches_torchvision():
    got = OI.to_tensor_normalize(G["img_u8"], (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    assert got.shape == G["img_normalized"].shape
    assert np.abs(got - G["img_normalized"]).max() <= 1e-6
