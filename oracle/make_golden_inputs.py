"""Generates tests/golden/inputs_golden.npz by running the UNMODIFIED reference in this container:
``DualEEGDataset._preprocess_eeg`` (imported by file path from /root/reference, called unbound -- it does not use
``self``) and the exact lines of its ``enable_preprocessing=False`` branch, plus torchvision's ToTensor + Normalize that
``multimodal_dataset.py:73-83`` composes.  Run once:  python oracle/make_golden_inputs.py"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/1_Data/processed/dual_eeg_dataset.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs_golden.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_dual_eeg_dataset", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(7)
    t = np.arange(384) / 256.0
    eeg = (rng.standard_normal((8, 384)) * rng.uniform(0.5, 30.0, (8, 1)) + 40.0 * np.sin(2 * np.pi * 10 * t)[None, :]
           + rng.uniform(-200, 200, (8, 1))).astype(np.float32)          # micro-volt-like scales, offsets, a shared rhythm
    pre = mod.DualEEGDataset._preprocess_eeg(None, eeg.copy())
    simple = (eeg - eeg.mean()) / (eeg.std() + 1e-8)                      # dual_eeg_dataset.py:196-198 verbatim expression
    from torchvision import transforms
    from PIL import Image
    img = rng.integers(0, 256, (20, 12, 3), dtype=np.uint8)
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    img_out = tf(Image.fromarray(img)).numpy()
    np.savez_compressed(OUT, eeg=eeg, eeg_preprocessed=pre.astype(np.float32), eeg_simple=simple.astype(np.float32),
                        img_u8=img, img_normalized=img_out.astype(np.float32))
    print("wrote", OUT, {k: v.shape for k, v in np.load(OUT).items()})


if __name__ == "__main__":
    main()
