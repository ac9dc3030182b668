"""GPU parity of the input-side kernels (egb_eeg_window_normalize, egb_image_u8_normalize) against the oracle and the
golden vectors of the unmodified reference, at the golden size and at the benchmark geometries (32 x 1024, 64 x 2048,
224 x 224), plus the size-independent properties.  Run on the B200 box:  pytest -m gpu"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200.inputs import normalize_eeg_windows, normalize_images_u8
from oracle import inputs as OI

DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "inputs_golden.npz"))


def test_eeg_golden(cuda_device):
    x = torch.from_numpy(G["eeg"]).to(DEV)
    assert np.abs(normalize_eeg_windows(x).cpu().numpy() - G["eeg_preprocessed"]).max() <= 2e-5
    assert np.abs(normalize_eeg_windows(x, enable_preprocessing=False).cpu().numpy() - G["eeg_simple"]).max() <= 2e-5


@pytest.mark.parametrize("shape", [(3, 32, 1024), (2, 64, 2048), (5, 7, 100)])
@pytest.mark.parametrize("pre", [True, False])
def test_eeg_vs_oracle(cuda_device, shape, pre):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(shape) * rng.uniform(1, 50, (shape[0], shape[1], 1)) + rng.uniform(-300, 300, (shape[0], shape[1], 1))
         ).astype(np.float32)
    got = normalize_eeg_windows(torch.from_numpy(x).to(DEV), enable_preprocessing=pre).cpu().numpy()
    fn = OI.preprocess_eeg if pre else OI.simple_normalize
    want = np.stack([fn(w) for w in x])
    assert np.abs(got - want).max() <= 5e-5
    if pre:      # size-independent property: every channel of every window ends with zero mean and unit variance
        assert np.abs(got.mean(axis=2)).max() <= 1e-4 and np.abs(got.std(axis=2) - 1).max() <= 1e-3
    else:        # ... and the whole window in the plain mode
        flat = got.reshape(shape[0], -1)
        assert np.abs(flat.mean(axis=1)).max() <= 1e-4 and np.abs(flat.std(axis=1) - 1).max() <= 1e-3


def test_image_golden_and_batch(cuda_device):
    u8 = torch.from_numpy(G["img_u8"]).to(DEV).unsqueeze(0)
    got = normalize_images_u8(u8)[0].cpu().numpy()
    assert np.abs(got - G["img_normalized"]).max() <= 2e-6
    rng = np.random.default_rng(4)
    batch = rng.integers(0, 256, (3, 224, 224, 3), dtype=np.uint8)
    got = normalize_images_u8(torch.from_numpy(batch).to(DEV)).cpu().numpy()
    want = np.stack([OI.to_tensor_normalize(b, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)) for b in batch])
    assert got.shape == (3, 3, 224, 224) and np.abs(got - want).max() <= 2e-6
    with pytest.raises(TypeError):
        normalize_images_u8(torch.from_numpy(batch))          # CPU tensor: there is no fallback
