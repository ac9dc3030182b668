import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import gaze_pair_batch
from oracle import vit as V
warnings.simplefilter("ignore")
DEV = "cuda:0"
name = "vit_small_patch16_224"
heads = V.VIT_VARIANTS[name][2]
sd = V.init_vit_state_dict(name, 6, 3, "backbone.", seed=31)
m = EarlyFusionViT(name, num_classes=3, pretrained=False, fusion_mode="concat")
m.load_state_dict(sd, strict=True)
m = m.to(DEV).eval()
a, b = gaze_pair_batch(4, seed=32)
for labels in (torch.tensor([0, 1, 2, 1]), torch.tensor([0, 1, 2, 0])):
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    F.cross_entropy(V.early_fusion_forward(sdr, a, b, heads, "concat"), labels).backward()
    for pre_fp32 in (False, True, False):
        m.zero_grad(set_to_none=True)
        if pre_fp32:
            with precision("fp32"):
                F.cross_entropy(m(a.to(DEV), b.to(DEV)), labels.to(DEV)).backward()
            m.zero_grad(set_to_none=True)
        with precision("bf16"):
            F.cross_entropy(m(a.to(DEV), b.to(DEV)).float(), labels.to(DEV)).backward()
        rows = []
        for k, p in m.named_parameters():
            r = sdr[k].grad
            if r.abs().max() < 1e-7: continue
            g = p.grad.float().cpu()
            rows.append(((g - r).abs().max().item() / r.abs().max().item(), (g - r).norm().item() / r.norm().item(), k))
        rows.sort(reverse=True)
        print(labels.tolist(), "pre_fp32", pre_fp32, ["%s %.3e %.3e" % (k, e, f) for e, f, k in rows[:3]])
