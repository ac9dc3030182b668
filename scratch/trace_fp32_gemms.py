"""Lists the fp32 (FFMA-route) GEMM calls of one cfg2 training step with their shapes and device times."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eyegaze_multimodal_b200 import ops, _lib as L
from eyegaze_multimodal_b200.multimodal import multimodal_loss
from eyegaze_multimodal_b200.precision import set_precision

dev = torch.device("cuda:0")
set_precision("bf16")
wl = bench.WORKLOADS["cfg2"]
model = bench.build_model(wl, dev).train()
B = 256
g = torch.Generator().manual_seed(0)
batch = dict(img1=torch.randn(B, 3, 224, 224, generator=g).to(dev), img2=torch.randn(B, 3, 224, 224, generator=g).to(dev),
             eeg1=torch.randn(B, 32, 1024, generator=g).to(dev), eeg2=torch.randn(B, 32, 1024, generator=g).to(dev),
             labels=torch.randint(0, 3, (B,), generator=g).to(dev))
model.concurrent_branches = False
log = []
orig = ops.gemm


def traced(M, N, K, in_code, a, b, c, **kw):
    if in_code == L.F32 or N < 16:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(M, N, K, in_code, a, b, c, **kw); e1.record()
        log.append((M, N, K, in_code, kw.get("accumulate", 0), kw.get("act", 0), e0, e1))
    else:
        orig(M, N, K, in_code, a, b, c, **kw)


def step():
    model.zero_grad(set_to_none=True)
    out = model(batch["img1"], batch["img2"], batch["eeg1"], batch["eeg2"], batch["labels"])
    multimodal_loss(model, out, batch["labels"]).backward()


for _ in range(2):
    step()
ops.gemm = traced
step()
torch.cuda.synchronize()
tot = 0.0
for M, N, K, code, acc, act, e0, e1 in log:
    us = e0.elapsed_time(e1) * 1e3
    tot += us
    print(f"M={M:6d} N={N:5d} K={K:6d} in={'f32' if code == L.F32 else 'bf16'} accumulate={acc} act={act}  {us:7.1f} us")
print("total %.1f us in %d calls" % (tot, len(log)))
