"""Times the attention backward (ViT-B and EEG shapes, incl. dropout and cross-attention lengths) for the variant
selected by EGB_ATT_FUSED_BWD and prints checksums so variants can be compared across processes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
for (S, Lq, D, H, pdrop, tag) in [(256, 197, 768, 12, 0.0, "ViT-B"), (512, 139, 256, 8, 0.1, "EEG"), (512, 139, 256, 8, 0.0, "EEG-nodrop"),
                                  (64, 33, 256, 8, 0.0, "L33"), (64, 235, 256, 8, 0.1, "L235"), (32, 256, 768, 12, 0.0, "L256d64"),
                                  (32, 64, 768, 12, 0.0, "L64d64"), (32, 130, 768, 12, 0.0, "L130d64")]:
    qkv = (torch.randn(S, Lq, 3 * D, device=dev) * 0.5).bfloat16().requires_grad_(True)
    go = torch.randn(S, Lq, D, device=dev).bfloat16()
    ts = []
    for it in range(6):
        qkv.grad = None
        o = ops.attention_packed(qkv, H, p=pdrop)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o.backward(go); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    g = qkv.grad.float()
    print("%-10s bwd %.1f us  |dqkv| %.6e  sum %.6e  finite %s" % (tag, min(ts), g.norm().item(), g.sum().item(), bool(torch.isfinite(g).all())))
