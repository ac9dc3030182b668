"""Attention kernels alone at the bench's shapes: CUDA-event timing (default) or a single pass for ncu (--once).

    python tools/att_bench.py [--once] [--iters 20]
ViT-B layer: S=256, L=197, 12 heads x 64;  ViT-S: 6 heads x 64;  EEG encoder layer: S=512, L=139, 8 heads x 32, dropout
0.1;  cfg5 EEG layer: S=1024, L=235."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eyegaze_multimodal_b200 import ops  # noqa: E402

SHAPES = [("vit_b", 256, 197, 768, 12, 0.0), ("vit_s", 256, 197, 384, 6, 0.0), ("eeg", 512, 139, 256, 8, 0.1),
          ("eeg_nodrop", 512, 139, 256, 8, 0.0), ("eeg_cfg5", 1024, 235, 256, 8, 0.1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for tag, S, Lq, D, H, p in SHAPES:
        if a.only and tag not in a.only.split(","):
            continue
        qkv = (torch.randn(S, Lq, 3 * D, device=dev) * 0.5).bfloat16().requires_grad_(True)
        go = torch.randn(S, Lq, D, device=dev).bfloat16()
        if a.once:
            ops.attention_packed(qkv, H, p=p).backward(go)
            torch.cuda.synchronize()
            continue
        for _ in range(3):
            ops.attention_packed(qkv, H, p=p).backward(go)
        tf = tb = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            o = ops.attention_packed(qkv, H, p=p)
            e[1].record()
            o.backward(go)
            e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1])
            tb += e[1].elapsed_time(e[2])
        dk = D // H
        fl = 4.0 * S * H * Lq * Lq * dk
        byt = 2.0 * S * Lq * D
        print("%-10s fwd %7.1f us (%6.1f TFLOP/s, %5.0f GB/s of q,k,v,o)   bwd %7.1f us (%6.1f TFLOP/s, %5.0f GB/s of 8 tensors)"
              % (tag, tf / a.iters * 1e3, fl / (tf / a.iters * 1e-3) / 1e12, 4 * byt / (tf / a.iters * 1e-3) / 1e9,
                 tb / a.iters * 1e3, 2.5 * fl / (tb / a.iters * 1e-3) / 1e12, 8 * byt / (tb / a.iters * 1e-3) / 1e9))


if __name__ == "__main__":
    main()
