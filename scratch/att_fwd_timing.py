"""Phase timers (clock64 stamps of a mid-grid CTA) and launch time of the attention forward kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200 import _lib as L, ops
dev = "cuda:0"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
names = ["issue_loads", "wait_loads+sync", "S_mma", "softmax(p1+p2)", "PV_mma", "store_O", "dealloc"]
torch.manual_seed(0)
for (S, Lq, D, H, pdrop, tag) in [(256, 197, 768, 12, 0.0, "ViT-B"), (512, 139, 256, 8, 0.1, "EEG"), (512, 139, 256, 8, 0.0, "EEG-nodrop")]:
    qkv = (torch.randn(S, Lq, 3 * D, device=dev) * 0.5).bfloat16()
    ts = []
    with torch.no_grad():
        for it in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if it == 4:
                L.call("egb_debug_attention_timing", buf.data_ptr())
            e0.record(); o = ops.attention_packed(qkv, H, p=pdrop); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        L.call("egb_debug_attention_timing", None)
    t = buf.cpu().tolist()
    print("%-10s fwd %.1f us  sum %.6e |" % (tag, min(ts), o.float().sum().item()), " ".join("%s=%d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)), "total", t[7] - t[0])
