"""How long the host needs to ENQUEUE one cfg2 training step (forward + loss + backward), measured with an empty GPU
queue at the start and no synchronisation at the end: the margin the launch path has against the 53 ms the GPU needs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eyegaze_multimodal_b200.multimodal import multimodal_loss
from eyegaze_multimodal_b200.parallel import TrialParallel
from eyegaze_multimodal_b200.precision import set_precision

dev = torch.device("cuda:0")
set_precision("bf16")
wl = bench.WORKLOADS["cfg2"]
model = bench.build_model(wl, dev).train()
tp = TrialParallel(model)
B = 256
g = torch.Generator().manual_seed(0)
b = dict(img1=torch.randn(B, 3, 224, 224, generator=g).to(dev), img2=torch.randn(B, 3, 224, 224, generator=g).to(dev),
         eeg1=torch.randn(B, 32, 1024, generator=g).to(dev), eeg2=torch.randn(B, 32, 1024, generator=g).to(dev),
         labels=torch.randint(0, 3, (B,), generator=g).to(dev))


def step():
    tp.zero_grad()
    out = tp(b["img1"], b["img2"], b["eeg1"], b["eeg2"], b["labels"])
    multimodal_loss(model, out, b["labels"]).backward()
    tp.finish()


for _ in range(3):
    step()
host, total = [], []
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host.append((t1 - t0) * 1e3)
    total.append((t2 - t0) * 1e3)
print(json.dumps({"host_enqueue_ms_per_step": sorted(host)[len(host) // 2], "step_ms_from_empty_queue": sorted(total)[len(total) // 2],
                  "host_cores": os.cpu_count()}))
