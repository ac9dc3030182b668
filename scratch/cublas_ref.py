"""cuBLAS (torch.matmul, bf16) on the hot GEMM shapes of cfg2 next to our kernels: what a tuned library reaches here."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200 import ops
dev = "cuda:0"
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = [("vit_fc2", 50432, 768, 3072), ("vit_qkv", 50432, 2304, 768), ("vit_fc1", 50432, 3072, 768), ("vit_proj", 50432, 768, 768),
          ("eeg_ffn1", 71168, 1024, 256), ("eeg_qkv", 71168, 768, 256), ("cube8k", 8192, 8192, 8192)]
for name, M, N, K in shapes:
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = torch.nn.Parameter(torch.randn(N, K, device=dev) / K ** 0.5)
    wb = w.detach().bfloat16()
    b = torch.nn.Parameter(torch.zeros(N, device=dev))
    bb = b.detach().bfloat16()
    with torch.no_grad():
        t_lib = bench(lambda: torch.nn.functional.linear(x, wb, bb))
        t_ours = bench(lambda: ops.linear(x, w, b))
        # dW: dy^T x
        dy = torch.randn(M, N, device=dev).bfloat16()
        t_lib_dw = bench(lambda: dy.t() @ x)
        t_ours_dw = bench(lambda: ops._grad_weight(dy, x, N, K))
    f = 2.0 * M * N * K * 1e-9
    print(f"{name:9s} M={M} N={N} K={K}: y=xW^T+b cuBLAS {t_lib*1e3:7.1f} us {f/t_lib:7.1f} TF/s | ours {t_ours*1e3:7.1f} us {f/t_ours:7.1f} TF/s || dW cuBLAS(bf16 out) {t_lib_dw*1e3:7.1f} us {f/t_lib_dw:7.1f} | ours(fp32 split-K) {t_ours_dw*1e3:7.1f} us {f/t_ours_dw:7.1f}")
