"""Input-side kernels at the cfg2 geometry (256 trials: 2 x 32 x 1024 EEG windows, 2 x 224 x 224 uint8 images per trial):
time and achieved HBM bandwidth (algorithmic bytes: EEG read + write 8 B / sample; image 3 B read + 12 B written / pixel)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200.inputs import normalize_eeg_windows, normalize_images_u8
dev = "cuda:0"
B = 512
eeg = torch.randn(B, 32, 1024, device=dev) * 20 + 5
img = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {}
for name, fn, nbytes in (("eeg_car_zscore", lambda: normalize_eeg_windows(eeg), eeg.numel() * 8),
                         ("eeg_window_zscore", lambda: normalize_eeg_windows(eeg, False), eeg.numel() * 8),
                         ("image_u8_normalize", lambda: normalize_images_u8(img), img.numel() * 5)):
    ms = timeit(fn)
    out[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "algorithmic_bytes": nbytes}
print(json.dumps(out))
