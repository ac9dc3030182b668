import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200 import _lib as L, ops
dev = "cuda:0"
buf = torch.zeros(8, dtype=torch.int64, device=dev)
names = ["s0", "s1", "s2", "s3", "s4", "s5", "s6"]
for (S, Lq, D, H, tag) in [(256, 197, 768, 12, "ViT-B"), (512, 139, 256, 8, "EEG")]:
    qkv = (torch.randn(S, Lq, 3 * D, device=dev) * 0.5).bfloat16().requires_grad_(True)
    go = torch.randn(S, Lq, D, device=dev).bfloat16()
    for it in range(3):
        o = ops.attention_packed(qkv, H)
        o.backward(go)
    torch.cuda.synchronize()
    for which in ("fwd", "bwd"):
        L.call("egb_debug_attention_timing", buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if which == "fwd":
            e0.record(); o = ops.attention_packed(qkv, H); e1.record()
            torch.cuda.synchronize()
            t = buf.cpu().tolist()
            print(tag, "fwd  %.1f us" % (e0.elapsed_time(e1) * 1e3), " ".join("%s=%d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)), "total", t[7] - t[0])
        else:
            o = ops.attention_packed(qkv, H)
            torch.cuda.synchronize()
            e0.record(); o.backward(go); e1.record()
            torch.cuda.synchronize()
            t = buf.cpu().tolist()   # last kernel = dkv
            print(tag, "bwd  %.1f us (dq+dkv); dkv CTA:" % (e0.elapsed_time(e1) * 1e3), " ".join("%s=%d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)), "total", t[7] - t[0])
    L.call("egb_debug_attention_timing", None)
