"""Per-shape table of the tensor-core GEMM launches of one bench step (CUDA events around every launch, encoder branches
serialised): launches, average us, TFLOP/s, fraction of the measured sustained bf16 peak.
    python tools/gemm_table.py [--workload cfg2]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from eyegaze_multimodal_b200 import _lib as L  # noqa: E402
from eyegaze_multimodal_b200.multimodal import multimodal_loss  # noqa: E402
from eyegaze_multimodal_b200.precision import set_precision  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--batch", type=int, default=0)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
B = a.batch or wl["batch"]
dev = torch.device("cuda:0")
set_precision("bf16")
model = bench.build_model(wl, dev).train()
g = torch.Generator().manual_seed(0)
e1 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
e2 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
lab = torch.randint(0, 3, (B,), generator=g).to(dev)
mm = wl["vit"] is not None
if mm:
    i1 = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    i2 = torch.randn(B, 3, 224, 224, generator=g).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    if mm:
        out = model(i1, i2, e1, e2, lab)
        loss = multimodal_loss(model, out, lab)
    else:
        out = model(e1, e2, lab)
        loss = out["loss"] + out.get("loss_ibs_cls", 0)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
if getattr(model, "concurrent_branches", False):
    model.concurrent_branches = False          # one kernel on the device at a time, as under ncu
for _ in range(2):
    step()
torch.cuda.synchronize()
L.prof_read(0, reset=True)
L.prof_enable(True)
N = 2
for _ in range(N):
    step()
torch.cuda.synchronize()
L.prof_enable(False)
recs = L.prof_dump(0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in recs:
    a_ = agg[r["tag"]]
    a_[0] += 1
    a_[1] += r["ms"]
    a_[2] += r["flops"]
tot_ms = sum(v[1] for v in agg.values())
print("%d launches / step, %.2f ms / step, %.0f TFLOP/s average (%.2f of %.0f)" % (
    len(recs) / N, tot_ms / N, sum(v[2] for v in agg.values()) / tot_ms / 1e9, sum(v[2] for v in agg.values()) / tot_ms / 1e9 / peak, peak))
print("%8s %6s %6s  %-22s %4s %8s %8s %7s %6s   (kernel<BN, epilogue mask>)" % ("M", "N", "K", "kernel", "n", "us", "ms/step", "TFLOP/s", "frac"))
for tag, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M_, N_, K_, var = tag
    kind = {1: "single", 2: "pair/8w", 3: "pair/16w", 4: "conv3x3"}.get(int(var // 1e10), "?")
    kern = "%s<%d,%d>" % (kind, int(var % 1e10) // 10000000, int(var % 1e7))
    print("%8d %6d %6d  %-22s %4d %8.1f %8.3f %7.0f %6.2f" % (M_, N_, K_, kern, n // N, ms / n * 1e3, ms / N, fl / ms / 1e9, fl / ms / 1e9 / peak))
