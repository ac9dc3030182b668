"""fp32 gradients of the default EEG model (32 ch x 1024, B = 4) against the CPU oracle in fp32 AND float64: is a deviation a
defect of a kernel variant or the conditioning of the gradient itself?  Usage: python tools/fp32_grad_ab.py [out.pt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import eeg_pair_batch
from oracle import eeg as O

DEV = "cuda:0"
cfg = O.EEGConfig(in_channels=32, max_len=256)
B, T, seed = 4, 1024, 2
sd = O.init_state_dict(cfg, seed)
m = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
m.load_state_dict(sd, strict=True)
m = m.to(DEV).eval()
e1, e2 = eeg_pair_batch(B, cfg.in_channels, T, seed=seed, coupled=True)
labels = torch.arange(B) % 3


def oracle(dt):
    sdr = {k: (v.clone().to(dt) if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
    for v in sdr.values():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
    ref = O.dual_eeg_forward(sdr, e1.to(dt), e2.to(dt), cfg, labels)
    (ref["loss"] + ref.get("loss_ibs_cls", 0.0)).backward()
    return {k: v.grad for k, v in sdr.items() if v.dtype.is_floating_point and v.grad is not None}


g32 = oracle(torch.float32)
try:
    g64 = oracle(torch.float64)
except Exception as ex:  # noqa: BLE001
    print("float64 oracle not available:", repr(ex)[:200])
    g64 = None
with precision("fp32"):
    out = m(e1.to(DEV), e2.to(DEV), labels.to(DEV))
    (out["loss"] + out.get("loss_ibs_cls", 0.0)).backward()
gg = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
if len(sys.argv) > 1:
    torch.save(gg, sys.argv[1])
print("%-52s %10s %12s %12s %12s" % ("parameter", "max|ref|", "gpu-cpu32", "gpu-cpu64", "cpu32-cpu64"))
for k in gg:
    if k not in g32:
        continue
    r = g32[k]
    mx = r.abs().max().item()
    a = (gg[k] - r).abs().max().item() / (mx + 1e-30)
    b = c = float("nan")
    if g64 is not None and k in g64:
        b = (gg[k].double() - g64[k]).abs().max().item() / (mx + 1e-30)
        c = (r.double() - g64[k]).abs().max().item() / (mx + 1e-30)
    if a > 5e-4 or b > 5e-4 or c > 5e-4:
        print("%-52s %10.3e %12.3e %12.3e %12.3e" % (k, mx, a, b, c))
