"""Kernel-level time table of one bench step via torch.profiler (CUPTI); iteration aid, not a bench number."""
import sys, os, collections, re, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eyegaze_multimodal_b200.multimodal import multimodal_loss
from eyegaze_multimodal_b200.precision import set_precision

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
B = a.batch or wl["batch"]
dev = torch.device("cuda:0")
set_precision("bf16")
model = bench.build_model(wl, dev).train()
g = torch.Generator().manual_seed(0)
e1 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev); e2 = torch.randn(B, wl["C"], wl["T"], generator=g).to(dev)
lab = torch.randint(0, 3, (B,), generator=g).to(dev)
mm = wl["vit"] is not None
if mm:
    i1 = torch.randn(B, 3, 224, 224, generator=g).to(dev); i2 = torch.randn(B, 3, 224, 224, generator=g).to(dev)
def step():
    model.zero_grad(set_to_none=True)
    if mm:
        out = model(i1, i2, e1, e2, lab); loss = multimodal_loss(model, out, lab)
    else:
        out = model(e1, e2, lab); loss = out["loss"] + out.get("loss_ibs_cls", 0)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
N = 2
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(N): step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "")).replace("void ", "")
        agg[n][0] += 1; agg[n][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"GPU busy {tot/N/1e3:.2f} ms/step over {N} steps")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"{t/N/1e3:8.3f} ms {100*t/tot:5.1f}% {c//N:5d} x {t/c:9.1f} us  {n[:90]}")
