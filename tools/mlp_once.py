"""One forward + backward of the ViT-B MLP block (fc1 + GELU -> fc2 + residual) at the bench's shape, for ncu:
    ncu --set full --import-source on -k regex:gemm_tc2 -c 8 python tools/mlp_once.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eyegaze_multimodal_b200 import _lib as L  # noqa: E402
from eyegaze_multimodal_b200 import ops  # noqa: E402

dev = "cuda:0"
EEG = "--eeg" in sys.argv            # the EEG encoder's FFN: ReLU, dropout 0.1 on both layers, K = 256
M, D, H = (71168, 256, 1024) if EEG else (50432, 768, 3072)
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(M, D, device=dev, generator=g) * 0.5).bfloat16().requires_grad_(True)
w1 = torch.nn.Parameter(torch.randn(H, D, device=dev, generator=g) / D ** 0.5)
b1 = torch.nn.Parameter(torch.zeros(H, device=dev))
w2 = torch.nn.Parameter(torch.randn(D, H, device=dev, generator=g) / H ** 0.5)
b2 = torch.nn.Parameter(torch.zeros(D, device=dev))
go = torch.randn(M, D, device=dev, generator=g).bfloat16()
ops.enable_seed_epoch()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    if EEG:
        y = ops.mlp2(x, w1, b1, w2, b2, L.ACT_RELU, p_mid=0.1, p_out=0.1, residual=x)
    else:
        y = ops.mlp2(x, w1, b1, w2, b2, L.ACT_GELU, residual=x)
    y.backward(go)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
