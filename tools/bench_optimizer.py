"""Training-step tail on the cfg2 model's parameters: FusedClipAdamW (2 launches) vs clip_grad_norm_ + torch.optim.AdamW
(foreach and fused=True).  HBM roofline: 16 B read + 12 B written per parameter for the update, + 4 B for the norm."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eyegaze_multimodal_b200.optim import FusedClipAdamW

dev = torch.device("cuda:0")
model = bench.build_model(bench.WORKLOADS["cfg2"], dev)
params = [p for p in model.parameters() if p.requires_grad]
n = sum(p.numel() for p in params)
for p in params:
    p.grad = torch.randn_like(p) * 1e-2


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


mine = FusedClipAdamW(params, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
t_mine = timeit(mine.step)
out = {"parameters": n, "tensors": len(params), "fused_clip_adamw_ms": t_mine,
       "fused_GBps": n * 32 / t_mine / 1e6, "hbm_bytes_per_param": 32}
for name, kw in (("torch_foreach", dict(foreach=True)), ("torch_fused", dict(fused=True))):
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, **kw)

    def ref_step():
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
    out[name + "_clip_adamw_ms"] = timeit(ref_step)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
out["peaks"] = {k: v for k, v in peaks.items() if "hbm" in k.lower()}
print(json.dumps(out))
