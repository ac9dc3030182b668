"""In-kernel phase timers of the pipelined attention backward (build with EGB_NVCC_DEFS=-DEGB_ATT_TIMING)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200 import _lib as L, ops
dev = "cuda:0"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
for (S, Lq, D, H, pdrop, tag) in [(256, 197, 768, 12, 0.0, "ViT-B"), (512, 139, 256, 8, 0.1, "EEG")]:
    qkv = (torch.randn(S, Lq, 3 * D, device=dev) * 0.5).bfloat16().requires_grad_(True)
    go = torch.randn(S, Lq, D, device=dev).bfloat16()
    for it in range(3):
        o = ops.attention_packed(qkv, H, p=pdrop)
        o.backward(go)
    torch.cuda.synchronize()
    o = ops.attention_packed(qkv, H, p=pdrop)
    torch.cuda.synchronize()
    L.call("egb_debug_attention_timing", buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o.backward(go); e1.record()
    torch.cuda.synchronize()
    L.call("egb_debug_attention_timing", None)
    t = buf.cpu().tolist()
    print(tag, "bwd %.1f us" % (e0.elapsed_time(e1) * 1e3),
          "issue_loads+stats=%d wait_loads+sync=%d rounds=%d final_dq=%d dkdv_store=%d sync=%d total=%d" %
          (t[1] - t[0], t[2] - t[1], t[3] - t[2], 0, t[4] - t[3], t[5] - t[4], t[5] - t[0]),
          "| in rounds: wait_sp=%d wait_acc=%d wait_dq(all)=%d elementwise=%d write+arrive=%d" % (t[8], t[9], t[10], t[11], t[12]))
