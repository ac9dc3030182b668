"""Shapes of the fp32 (FFMA) GEMM launches of one training step of a bench workload, in launch order, with event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eyegaze_multimodal_b200 import ops, _lib as L
from eyegaze_multimodal_b200.multimodal import multimodal_loss
from eyegaze_multimodal_b200.precision import set_precision

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B = wl["batch"]
dev = torch.device("cuda:0")
set_precision("bf16")
model = bench.build_model(wl, dev).train()
if getattr(model, "concurrent_branches", False):
    model.concurrent_branches = False
g = torch.Generator().manual_seed(0)
e1 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
e2 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
lab = torch.randint(0, 3, (B,), generator=g).to(dev)
i1 = torch.randn(B, 3, 224, 224, generator=g).to(dev)
i2 = torch.randn(B, 3, 224, 224, generator=g).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    out = model(i1, i2, e1, e2, lab)
    multimodal_loss(model, out, lab).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
log = []
orig = ops.gemm


def spy(M, N, K, in_code, a, b, c, **kw):
    if in_code == L.F32:
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(M, N, K, in_code, a, b, c, **kw)
        e1_.record()
        log.append((M, N, K, a[1], b[1], kw.get("act", 0), kw.get("accumulate", 0), e0, e1_))
    else:
        orig(M, N, K, in_code, a, b, c, **kw)


ops.gemm = spy
step()
torch.cuda.synchronize()
tot = 0.0
for M, N, K, am, bm, act, acc, e0, e1_ in log:
    t = e0.elapsed_time(e1_) * 1e3
    tot += t
    print("M=%6d N=%5d K=%6d a_major=%d b_major=%d act=%d acc=%d  %7.1f us  %6.1f GFLOP/s" % (M, N, K, am, bm, act, acc, t, 2.0 * M * N * K / t / 1e3))
print("%d launches, %.1f us" % (len(log), tot))
