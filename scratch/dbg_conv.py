import math, torch, torch.nn.functional as F, sys
sys.path.insert(0, '.')
from eyegaze_multimodal_b200 import ops, _lib as L
DEV='cuda:0'
torch.manual_seed(6)
for (B,C,T,D) in [(3,8,256,32),(2,32,1024,256)]:
    e1, e2 = torch.randn(B, C, T), torch.randn(B, C, T)
    w1, b1 = torch.randn(D, C, 25) / math.sqrt(25 * C), torch.randn(D) * 0.1
    w2, b2 = torch.randn(D, D, 25) / math.sqrt(25 * D), torch.randn(D) * 0.1
    ps = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    def ref(x):
        h1 = F.relu(F.conv1d(x, ps[0], ps[1], stride=4, padding=12)); h1.retain_grad()
        h = F.relu(F.conv1d(h1, ps[2], ps[3], stride=4, padding=12))
        return h.permute(0, 2, 1), h1
    r1, h1a = ref(e1); r2, h1b = ref(e2)
    hr = torch.cat([r1, r2], 0)
    gh = torch.randn_like(hr)
    hr.backward(gh)
    prm = [torch.nn.Parameter(t.clone().to(DEV)) for t in (w1, b1, w2, b2)]
    h = ops.temporal_conv(e1.to(DEV), e2.to(DEV), [prm[0], prm[2]], [prm[1], prm[3]], L.F32, 4, 0.0)
    h.backward(gh.to(DEV))
    print("cfg", (B,C,T,D), "h err", (h.cpu()-hr).abs().max().item())
    for g, r, nm in zip(prm, ps, ["dw1", "db1", "dw2", "db2"]):
        print("  ", nm, "err", (g.grad.cpu()-r.grad).abs().max().item(), "ref max", r.grad.abs().max().item())
    # per-tap error of dw1
    e = (prm[0].grad.cpu()-ps[0].grad).abs().amax(dim=(0,1))
    print("   dw1 err per tap", [round(x,4) for x in e.tolist()])
    e = (prm[2].grad.cpu()-ps[2].grad).abs().amax(dim=(0,1))
    print("   dw2 err per tap", [round(x,4) for x in e.tolist()])
