"""Attention kernels alone at the bench's shapes, timed with CUDA events around CUDA-graph replays (the eager call path
costs ~100 us of host time per call -- more than the forward kernel runs -- so eager event pairs would time the host).

    python tools/att_bench.py [--once] [--iters 20] [--only vit_b,eeg]
ViT-B layer: S=256, L=197, 12 heads x 64;  ViT-S: 6 heads x 64;  EEG encoder layer: S=512, L=139, 8 heads x 32, dropout
0.1;  cfg5 EEG layer: S=1024, L=235.  --once: one eager forward + backward per shape (for ncu)."""
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
    ops.enable_seed_epoch()
    for tag, S, Lq, D, H, p in SHAPES:
        if a.only and tag not in a.only.split(","):
            continue
        qkv = (torch.randn(S, Lq, 3 * D, device=dev) * 0.5).bfloat16().requires_grad_(True)
        go = torch.randn(S, Lq, D, device=dev).bfloat16()
        if a.once:
            ops.attention_packed(qkv, H, p=p).backward(go)
            torch.cuda.synchronize()
            continue
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                qkv.grad = None
                ops.attention_packed(qkv, H, p=p).backward(go)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        qkv.grad = None
        ops.reset_arenas()
        gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(gf):
            o = ops.attention_packed(qkv, H, p=p)
        with torch.cuda.graph(gb, pool=gf.pool()):
            o.backward(go)
        ops.reset_arenas()
        tf = tb = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            gf.replay()
            e[1].record()
            gb.replay()
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
