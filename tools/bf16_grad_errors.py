"""Diagnostic: per-parameter bf16 gradient error (max-abs-err / max-abs-ref, and Frobenius-relative) of the CUDA path
against the fp32 CPU oracle, for the gaze backbones and the EEG model at several batch sizes."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer  # noqa: E402
from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT  # noqa: E402
from eyegaze_multimodal_b200.precision import precision  # noqa: E402
from eyegaze_multimodal_b200.synth import eeg_pair_batch, gaze_pair_batch  # noqa: E402
from oracle import eeg as O  # noqa: E402
from oracle import vit as V  # noqa: E402

DEV = "cuda:0"
warnings.simplefilter("ignore")


def report(tag, named, ref):
    rows = []
    for k, p in named:
        r = ref[k].grad
        if r is None or r.abs().max().item() < 1e-7:      # (k_proj.bias: analytically zero, softmax is shift invariant)
            continue
        g = p.grad.float().cpu()
        rows.append(((g - r).abs().max().item() / (r.abs().max().item() + 1e-30), (g - r).norm().item() / (r.norm().item() + 1e-30), k))
    rows.sort(reverse=True)
    print(tag, "worst max-rel %.3e  worst fro-rel %.3e" % (rows[0][0], max(r[1] for r in rows)))
    for e, f, k in rows[:6]:
        print("    %-50s max-rel %.3e fro-rel %.3e" % (k, e, f))


for name, Bs in (("vit_small_patch16_224", (4, 16)), ("vit_base_patch16_224", (4, 8))) if "--eeg-only" not in sys.argv else ():
    heads = V.VIT_VARIANTS[name][2]
    sd = V.init_vit_state_dict(name, 6, 3, "backbone.", seed=31)
    m = EarlyFusionViT(name, num_classes=3, pretrained=False, fusion_mode="concat")
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    for B in Bs:
        a, b = gaze_pair_batch(B, seed=32)
        labels = torch.arange(B) % 3
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        F.cross_entropy(V.early_fusion_forward(sdr, a, b, heads, "concat"), labels).backward()
        m.zero_grad(set_to_none=True)
        with precision("bf16"):
            F.cross_entropy(m(a.to(DEV), b.to(DEV)).float(), labels.to(DEV)).backward()
        report("%s B=%d" % (name, B), list(m.named_parameters()), sdr)

for cfg, T, Bs in ((O.EEGConfig(in_channels=32, max_len=128, use_spectrogram=False, use_ibs=False), 512, (8, 32)),
                   (O.EEGConfig(in_channels=32, max_len=256), 1024, (8,))):
    sd = O.init_state_dict(cfg, 2)
    m = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    for B in Bs:
        e1, e2 = eeg_pair_batch(B, cfg.in_channels, T, seed=2, coupled=True)
        labels = torch.arange(B) % 3
        sdr = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
        ref = O.dual_eeg_forward(sdr, e1, e2, cfg, labels)
        (ref["loss"] + ref.get("loss_ibs_cls", 0.0)).backward()
        m.zero_grad(set_to_none=True)
        with precision("bf16"):
            out = m(e1.to(DEV), e2.to(DEV), labels.to(DEV))
            (out["loss"] + out.get("loss_ibs_cls", 0.0)).backward()
        report("EEG use_ibs=%s T=%d B=%d" % (cfg.use_ibs, T, B), list(m.named_parameters()), sdr)
        # the same op chain in plain PyTorch on the GPU under bf16 autocast (ATen / cuBLAS kernels): how much of the error
        # above is the precision itself rather than this repository's kernels
        sdg = {k: v.to(DEV).clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o2 = O.dual_eeg_forward(sdg, e1.to(DEV), e2.to(DEV), cfg, labels.to(DEV))
            (o2["loss"] + o2.get("loss_ibs_cls", 0.0)).backward()

        class _P:
            def __init__(self, g):
                self.grad = g
        report("   torch autocast(bf16) on cuda, same inputs", [(k, _P(v.grad.cpu())) for k, v in sdg.items() if v.grad is not None], sdr)
        ratio = []
        for k, p in m.named_parameters():
            r = sdr[k].grad
            if r is None or r.abs().max().item() < 1e-7 or sdg[k].grad is None:
                continue
            eo = (p.grad.float().cpu() - r).norm().item() / r.norm().item()
            et = (sdg[k].grad.float().cpu() - r).norm().item() / r.norm().item()
            ratio.append((eo / max(et, 1e-9), eo, et, k))
        ratio.sort(reverse=True)
        print("   fro-rel ours / torch-autocast, per parameter (all):")
        for q, eo, et, k in ratio:
            print("      %-52s ours %.3e torch %.3e ratio %5.2f" % (k, eo, et, q))
        m.zero_grad(set_to_none=True)
        with precision("fp32"):
            out = m(e1.to(DEV), e2.to(DEV), labels.to(DEV))
            (out["loss"] + out.get("loss_ibs_cls", 0.0)).backward()
        report("   this path in fp32 mode", list(m.named_parameters()), sdr)
