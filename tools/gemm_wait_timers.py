import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200 import _lib as L, ops
dev = "cuda:0"
buf = torch.zeros(160, 8, dtype=torch.int64, device=dev)
def run(M, N, K, tag, bias=True, res=False):
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = torch.nn.Parameter(torch.randn(N, K, device=dev) / K ** 0.5)
    b = torch.nn.Parameter(torch.zeros(N, device=dev)) if bias else None
    r = torch.randn(M, N, device=dev).bfloat16() if res else None
    with torch.no_grad():
        for _ in range(3): ops.linear(x, w, b, residual=r)
        torch.cuda.synchronize()
        buf.zero_()
        L.call("egb_debug_gemm_timing", buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(x, w, b, residual=r); e1.record()
        torch.cuda.synchronize()
        L.call("egb_debug_gemm_timing", None)
    t = buf.float().cpu()
    act = t[t[:, 4] > 0]
    lead = act[act[:, 1] + act[:, 2] > 0]
    ms = e0.elapsed_time(e1)
    print(f"{tag:10s} M={M} N={N} K={K}: {ms*1e3:7.1f} us {2*M*N*K/ms*1e-9:7.1f} TF/s | kernel cyc {act[:,4].mean():9.0f} | producer wait-empty {act[:,0].mean()/act[:,4].mean():5.1%} | MMA wait-operands {lead[:,1].mean()/lead[:,4].mean():5.1%} wait-acc-drain {lead[:,2].mean()/lead[:,4].mean():5.1%} | epilogue wait-acc {act[:,3].mean()/act[:,4].mean():5.1%}")
print("EGB_GEMM_PAIR =", os.environ.get("EGB_GEMM_PAIR"), "EGB_GEMM_SKIPB =", os.environ.get("EGB_GEMM_SKIPB"))
run(50432, 768, 768, "proj_res", res=True)
run(50432, 768, 768, "proj_nores", res=False)
run(71168, 256, 256, "eeg_proj", res=True)
run(50432, 768, 3072, "fc2", res=True)
run(50432, 2304, 768, "qkv")
run(50432, 3072, 768, "fc1-nobias", bias=False)
run(71168, 1024, 256, "eeg_ffn1")
run(8192, 8192, 8192, "cublas-ref", bias=False)

def run_dw(M, N, K, tag):
    """dW[N_out=M, K_in=N] = dy^T x with K = rows (split-K accumulate)"""
    dy = torch.randn(K, M, device=dev).bfloat16()
    x = (torch.randn(K, N, device=dev) * 0.5).bfloat16()
    with torch.no_grad():
        for _ in range(3): ops._grad_weight(dy, x, M, N)
        torch.cuda.synchronize()
        buf.zero_()
        L.call("egb_debug_gemm_timing", buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops._grad_weight(dy, x, M, N); e1.record()
        torch.cuda.synchronize()
        L.call("egb_debug_gemm_timing", None)
    t = buf.float().cpu()
    act = t[t[:, 4] > 0]
    lead = act[act[:, 1] + act[:, 2] > 0]
    ms = e0.elapsed_time(e1)
    print(f"{tag:10s} dW M={M} N={N} K={K}: {ms*1e3:7.1f} us {2*M*N*K/ms*1e-9:7.1f} TF/s | ctas {len(act)} lead {len(lead)} kernel cyc {lead[:,4].mean():9.0f} (max {lead[:,4].max():9.0f} min {lead[:,4].min():9.0f}) | MMA wait-operands {lead[:,1].mean()/lead[:,4].mean():5.1%} wait-acc-drain {lead[:,2].mean()/lead[:,4].mean():5.1%}")
run_dw(3072, 768, 50432, "fc1_dw")
run_dw(2304, 768, 50432, "qkv_dw")
run_dw(768, 768, 50432, "proj_dw")
