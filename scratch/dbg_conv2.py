import math, torch, torch.nn.functional as F, sys
sys.path.insert(0, '.')
from eyegaze_multimodal_b200 import ops, _lib as L
DEV='cuda:0'
torch.manual_seed(6)
B,C,T,D = 3,8,256,32
e1, e2 = torch.randn(B, C, T), torch.randn(B, C, T)
w1, b1 = torch.randn(D, C, 25) / math.sqrt(25 * C), torch.randn(D) * 0.1
w2, b2 = torch.randn(D, D, 25) / math.sqrt(25 * D), torch.randn(D) * 0.1
ps = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
x = torch.cat([e1,e2],0)
pre1 = F.conv1d(x, ps[0], ps[1], stride=4, padding=12); pre1.retain_grad()
h1 = F.relu(pre1)
pre2 = F.conv1d(h1, ps[2], ps[3], stride=4, padding=12); pre2.retain_grad()
hr = F.relu(pre2).permute(0,2,1)
gh = torch.randn_like(hr)
hr.backward(gh)
prm = [torch.nn.Parameter(t.clone().to(DEV)) for t in (w1, b1, w2, b2)]
ops._debug = {}
h = ops.temporal_conv(e1.to(DEV), e2.to(DEV), [prm[0], prm[2]], [prm[1], prm[3]], L.F32, 4, 0.0)
print("h err", (h.cpu()-hr).abs().max().item())
h.backward(gh.to(DEV))
g0, front, rp = ops._debug["g0"]
T2 = hr.shape[1]
got = g0[:, front:front+T2].cpu()
want = pre2.grad.permute(0,2,1)
print("g0 shape", tuple(g0.shape), "front", front, "rp", rp, "dpre2 err", (got-want).abs().max().item(), "max", want.abs().max().item())
print("g0 pad nonzero", g0[:, :front].abs().max().item(), g0[:, front+T2:].abs().max().item())
print("gh vs got ratio sample", got[0,0,:6], want[0,0,:6], gh[0,0,:6])
g1, pad, tp = ops._debug["g1"]
T1 = pre1.shape[2]
got1 = g1[:, pad:pad+T1].cpu(); want1 = pre1.grad.permute(0,2,1)
print("dpre1 err", (got1-want1).abs().max().item(), "max", want1.abs().max().item())
