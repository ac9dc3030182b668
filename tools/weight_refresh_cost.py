"""Cost of re-deriving the bf16 / packed weight copies every step (what a real training step pays after the optimizer has
updated the fp32 masters): eager cfg2 steps with (a) cached copies, (b) one-launch refresh of the plain casts + lazy rest,
(c) everything lazy, one cast launch per parameter."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from eyegaze_multimodal_b200 import _lib as L  # noqa: E402
from eyegaze_multimodal_b200 import ops  # noqa: E402
from eyegaze_multimodal_b200.multimodal import multimodal_loss  # noqa: E402
from eyegaze_multimodal_b200.precision import set_precision  # noqa: E402

wl = bench.WORKLOADS["cfg2"]
B = wl["batch"]
dev = torch.device("cuda:0")
set_precision("bf16")
model = bench.build_model(wl, dev).train()
g = torch.Generator().manual_seed(0)
e1 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
e2 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
lab = torch.randint(0, 3, (B,), generator=g).to(dev)
i1 = torch.randn(B, 3, 224, 224, generator=g).to(dev)
i2 = torch.randn(B, 3, 224, 224, generator=g).to(dev)


def step(mode):
    if mode == "refresh":
        ops.bump_param_epoch()
        ops.refresh_plain_copies(dev)
    elif mode == "lazy":
        ops.bump_param_epoch()
    model.zero_grad(set_to_none=True)
    out = model(i1, i2, e1, e2, lab)
    multimodal_loss(model, out, lab).backward()


for mode in ("cached", "refresh", "lazy", "cached"):
    for _ in range(3):
        step(mode)
    torch.cuda.synchronize()
    n0 = L.launch_count()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step(mode)
    e1_.record()
    torch.cuda.synchronize()
    print("%-8s %.2f ms/step, %d launches/step" % (mode, e0.elapsed_time(e1_) / 10, (L.launch_count() - n0) // 10))
