#!/usr/bin/env python
"""bench.py -- train trials/sec (forward + backward) of the gaze+EEG fusion classifier on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference] [--mode train|infer]

One *step* = one pass of the hot path (MultimodalFusionModel forward, the 4-term loss of
train_multimodal_fuzzy_fusion.py:440-460, backward, and for N > 1 the gradient all-reduce) over one batch of
synthetic trials per GPU.  The step is captured ONCE into a CUDA graph (eyegaze_multimodal_b200/graphs.py) and replayed:
the graph re-derives the bf16 / packed weight copies from the fp32 master parameters on every replay and draws fresh
dropout masks (device seed epoch), i.e. it does everything an eager step does (--graphs 0 times the eager step).
Rank 0 prints ONE JSON line (see the contract in the task description):

  value     whole-job trials/s with the inputs already resident in HBM (device-timed with CUDA events,
            max over ranks);
  train_step  the same loop with the training-step tail inside: global-norm clip + AdamW (FusedClipAdamW) captured in
            the graph -- the reference's whole inner loop body (train_multimodal_fuzzy_fusion.py:425-504);
  eager     the un-captured step (launches issued from Python) and the host time it takes to enqueue it;
  gpu_eager_baseline  plain PyTorch-eager (ATen / cuBLAS / cuDNN kernels, bf16 autocast) on the same GPU, same workload:
            the "reference on the same B200" number of BASELINE.md 3.4;
  e2e       the same metric through the public nn.Module API with HOST inputs: every step copies its batch from
            pinned host memory to the device and reads the loss back, all inside the timed region;
  roofline  the dominant kernel (the tcgen05 bf16 GEMM): algorithmic FLOPs of its launches / their CUDA-event
            durations, against the measured sustained bf16 peak in MEASURED_PEAKS.json;
  cpu_baseline  the CPU oracle port (oracle/: the reference's algorithm restated in fp32 PyTorch-CPU) timed on this
            box's host cores on a bounded sample of the same workload (N = 1, rank 0 only).

``--impl reference`` times that CPU port alone, with all host threads, on the same workload / metric / unit
(the reference itself is pure Python + PyTorch and /root/reference does not exist on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train trials/sec (fwd+bwd)"
UNIT = "trials/s"

# name -> workload (BASELINE.json configs; SURVEY.md section 8d maps them to reference objects)
WORKLOADS = {
    # configs[1]: full gaze+EEG classifier as 4_Experiments/configs/multimodal_fuzzy_fusion.yaml builds it
    "cfg2": dict(desc="gaze+EEG fusion classifier (EarlyFusionViT vit_base_patch16_224 concat + DualEEGTransformer "
                      "32ch x 1024 + FuzzyGatingFusion full), batch 256 per GPU, fwd+bwd",
                 vit="vit_base_patch16_224", C=32, T=1024, batch=256, eeg_kwargs={}),
    # configs[3]: same with a ViT-S gaze backbone (data-parallel at 2/4/8 GPUs)
    "cfg4": dict(desc="cross-attention fusion with a ViT-S gaze-heatmap backbone, batch 256 per GPU, fwd+bwd",
                 vit="vit_small_patch16_224", C=32, T=1024, batch=256, eeg_kwargs={}),
    # configs[0]: EEG-only conv encoder + head (ablation A1: no spectrogram, no IBS), 32ch x 512, batch 64
    "cfg1": dict(desc="EEG-only temporal-conv encoder + head (A1_baseline_temporal_only), 32ch x 512, batch 64",
                 vit=None, C=32, T=512, batch=64, eeg_kwargs=dict(use_spectrogram=False, use_ibs=False)),
    # configs[4]: large sweep, 64ch x 2048, 4096 trials per step over 8 GPUs = 512 per GPU
    "cfg5": dict(desc="large sweep: 64ch x 2048 EEG, full fusion model (ViT-B), 512 trials per GPU",
                 vit="vit_base_patch16_224", C=64, T=2048, batch=512, eeg_kwargs={}),
    # configs[2]: the cross-modal attention unit alone: CrossBrainAttention(256, 8) on 64 query x 128 key tokens
    "cfg3": dict(desc="cross-modal attention unit (CrossBrainAttention d=256, 8 heads; 64 gaze tokens x 128 EEG tokens), "
                      "4096 trials per GPU, fwd+bwd", vit=None, C=0, T=0, batch=4096, eeg_kwargs={}, unit="cross_attention",
                 Lq=64, Lk=128, d=256, heads=8),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="trials per GPU (default: the workload's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="trials per CPU-baseline step")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer: forward only, eval mode")
    ap.add_argument("--graphs", type=int, default=1, help="1: replay the captured CUDA graph of the step; 0: eager launches")
    ap.add_argument("--no-train-step", action="store_true", help="skip the fwd+bwd+clip+AdamW leg")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager-on-GPU baseline leg")
    ap.add_argument("--bucket-mb", type=float, default=32.0, help="gradient all-reduce bucket size (N>1)")
    ap.add_argument("--grad-comm", default="fp32", choices=["fp32", "bf16"],
                    help="wire dtype of the gradient all-reduce (N>1); bf16 halves the bytes, the optimiser still sees fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# shared by both arms: the workload description the driver compares between them
# ------------------------------------------------------------------------------------------------------
def make_config(args, world):
    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    return {"workload": args.workload + ": " + wl["desc"], "trials_per_gpu": B, "global_batch": B * world,
            "parallelism": "dp%d (trial-wise shards, gradient all-reduce over NCCL%s)" % (
                world, ", %s on the wire" % args.grad_comm if world > 1 else ""),
            "mode": "train (dropout on), fwd+bwd" if args.mode == "train" else "inference (eval), fwd only",
            "l2": "inputs larger than L2: every step reads a fresh batch of synthetic trials"}


def metric_name(args):
    return METRIC if args.mode == "train" else "inference trials/sec (fwd)"


# ------------------------------------------------------------------------------------------------------
# baseline arms: the oracle port (fp32 PyTorch restatement of the reference, oracle/) -- on the host cores for the CPU
# baseline / --impl reference, and on the GPU under bf16 autocast for the PyTorch-eager-on-B200 baseline
# ------------------------------------------------------------------------------------------------------
def _oracle_step_fn(wl, B, device="cpu", train=True, seed=0):
    """Returns step(): one fwd (+bwd) of the workload on B synthetic trials through the oracle's plain PyTorch ops."""
    import torch
    from oracle import eeg as O
    from oracle import fuzzy as FZ
    from oracle import vit as V
    from eyegaze_multimodal_b200.synth import gaze_pair_batch, labels_batch, randn_eeg_pair

    def leaf(v, grad=True):
        v = v.clone().to(device)
        return v.requires_grad_(True) if (grad and train and v.is_floating_point()) else v
    if wl.get("unit") == "cross_attention":
        cfg = O.EEGConfig(in_channels=8, d_model=wl["d"], num_heads=wl["heads"], max_len=256)
        full = O.init_state_dict(cfg, seed=seed)
        sd = {k: leaf(v) for k, v in full.items() if k.startswith("cross_attn.")}
        g = torch.Generator().manual_seed(seed)
        z1 = torch.randn(B, wl["Lq"], wl["d"], generator=g).to(device)
        z2 = torch.randn(B, wl["Lk"], wl["d"], generator=g).to(device)

        def step():
            a, b = O.cross_brain(z1, z2, sd, cfg)
            loss = a.float().square().mean() + b.float().square().mean()
            if train:
                loss.backward()
            return loss
        return step
    cfg = O.EEGConfig(in_channels=wl["C"], max_len=wl["T"] // 4, **wl["eeg_kwargs"])
    sd = {k: leaf(v, grad=(k != "spectrogram_generator.window")) for k, v in O.init_state_dict(cfg, seed=seed).items()}
    e1, e2 = (t.to(device) for t in randn_eeg_pair(B, wl["C"], wl["T"], seed=seed))
    labels = labels_batch(B, seed=seed).to(device)
    if wl["vit"] is None:
        def step():
            out = O.dual_eeg_forward(sd, e1, e2, cfg, labels)
            loss = out["loss"] + out["loss_ibs_cls"] if "loss_ibs_cls" in out else out["loss"]
            if train:
                loss.backward()
            return loss
        return step
    heads = V.VIT_VARIANTS[wl["vit"]][2]
    vsd = {k: leaf(v) for k, v in V.init_vit_state_dict(wl["vit"], 6, 3, "backbone.", seed=seed).items()}
    fz = {k: leaf(v, grad=(k != "c_reliable")) for k, v in FZ.init_params().items()}
    a, b = (t.to(device) for t in gaze_pair_batch(B, seed=seed))

    def step():
        img_logits = V.early_fusion_forward(vsd, a, b, heads, "concat")
        eeg_logits = O.dual_eeg_forward(sd, e1, e2, cfg, labels)["logits"]
        fused, _alpha, aux = FZ.fuzzy_forward(fz, img_logits.float(), eeg_logits.float(), "full")
        loss = FZ.multimodal_loss(fused, img_logits.float(), eeg_logits.float(), aux, FZ.temperature_regularization(fz), labels)
        if train:
            loss.backward()
        return loss
    return step


def time_cpu(wl, B, steps, warmup, train=True):
    import torch
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    step = _oracle_step_fn(wl, B, "cpu", train)
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            float(step().detach())
        dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps, n_threads


def time_gpu_eager(wl, B, device, steps, warmup, train=True):
    """PyTorch-eager on the same GPU: the oracle's op chain on CUDA tensors under bf16 autocast --
    ATen / cuBLAS / cuDNN / cuFFT kernels, none of this repository's.  Batch is halved until it fits."""
    import torch
    while B >= 1:
        try:
            step = _oracle_step_fn(wl, B, device, train)
            ctx = torch.enable_grad() if train else torch.no_grad()
            with ctx, torch.autocast("cuda", dtype=torch.bfloat16):
                for _ in range(warmup):
                    step()
                torch.cuda.synchronize(device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1) / steps
            return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B,
                    "what": "oracle port (plain PyTorch ops) on cuda under torch.autocast(bfloat16), %d timed steps, "
                            "fwd%s; vectorised IBS restatement (the reference's own Python pair loops are ~20x slower)"
                            % (steps, "+bwd" if train else " only")}
        except torch.OutOfMemoryError:
            torch.cuda.empty_cache()
            B //= 2
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0                      # the CPU arm does not shard: rank 0 alone runs it
    wl = WORKLOADS[args.workload]
    B = args.cpu_sample
    train = args.mode == "train"
    tps, s_per_step, n_threads = time_cpu(wl, B, args.steps, args.warmup, train)
    sample = ("%d trials per step (a bounded sample of the workload's batch), fp32, all %d host threads, %s; the oracle "
              "port of the reference (its IBS generator restated in vectorised form: the reference's own Python pair loops "
              "are ~20x slower)" % (B, n_threads, "fwd+bwd" if train else "fwd only"))
    line = {
        "impl": "reference", "metric": metric_name(args), "value": tps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def build_model(wl, device):
    import warnings

    import torch
    from eyegaze_multimodal_b200.dual_eeg_transformer import CrossBrainAttention, DualEEGTransformer
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    from eyegaze_multimodal_b200.multimodal import MultimodalFusionModel

    torch.manual_seed(0)
    if wl.get("unit") == "cross_attention":
        return CrossBrainAttention(wl["d"], wl["heads"], dropout=0.1).to(device)      # dual_eeg_transformer.py:944-974
    # constructor calls of train_multimodal_fuzzy_fusion.py:653-700 / train_art.py:360-385
    eeg = DualEEGTransformer(in_channels=wl["C"], num_classes=3, d_model=256, num_layers=6, num_heads=8, d_ff=1024,
                             dropout=0.1, max_len=wl["T"] // 4, conv_kernel_size=25, conv_stride=4, conv_layers=2,
                             sampling_rate=256, **wl["eeg_kwargs"])
    if wl["vit"] is None:
        return eeg.to(device)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gaze = EarlyFusionViT(model_name=wl["vit"], num_classes=3, pretrained=False, img_size=224, fusion_mode="concat",
                              weight_init_strategy="duplicate")
    return MultimodalFusionModel(gaze, eeg, FuzzyGatingFusion(num_classes=3, mode="full", eps_temp=0.1)).to(device)


def host_batch(wl, B, g, pin=True):
    """One synthetic batch of the workload's shape on the host (pinned)."""
    import torch
    if wl.get("unit") == "cross_attention":
        hb = {"z1": torch.randn(B, wl["Lq"], wl["d"], generator=g), "z2": torch.randn(B, wl["Lk"], wl["d"], generator=g)}
    else:
        hb = {"eeg1": torch.randn(B, wl["C"], wl["T"], generator=g), "eeg2": torch.randn(B, wl["C"], wl["T"], generator=g),
              "labels": torch.randint(0, 3, (B,), generator=g)}
        if wl["vit"] is not None:
            hb["img1"] = torch.randn(B, 3, 224, 224, generator=g)
            hb["img2"] = torch.randn(B, 3, 224, 224, generator=g)
    return {k: v.pin_memory() for k, v in hb.items()} if pin else hb


def make_loss_fn(wl, tp, train):
    """loss_fn(model, batch) -> {'loss': scalar}: the forward + loss of one step through the drop-in modules."""
    from eyegaze_multimodal_b200.multimodal import multimodal_loss

    if wl.get("unit") == "cross_attention":
        def unit_loss(model, batch):
            a, b = tp(batch["z1"], batch["z2"])
            return {"loss": a.float().square().mean() + b.float().square().mean()}
        return unit_loss
    if wl["vit"] is not None:
        def mm_loss(model, batch):
            out = tp(batch["img1"], batch["img2"], batch["eeg1"], batch["eeg2"], batch["labels"])
            if not train:
                return {"loss": out["fused_logits"].float().sum(), "fused_logits": out["fused_logits"]}
            return {"loss": multimodal_loss(model, out, batch["labels"])}
        return mm_loss

    def eeg_loss(model, batch):
        out = tp(batch["eeg1"], batch["eeg2"], batch["labels"])
        return {"loss": out["loss"] + out["loss_ibs_cls"] if "loss_ibs_cls" in out else out["loss"]}
    return eeg_loss


def run_b200(args):
    import torch
    import torch.distributed as dist
    from eyegaze_multimodal_b200 import _lib as L
    from eyegaze_multimodal_b200 import ops
    from eyegaze_multimodal_b200.graphs import GraphedForward, GraphedTrainStep
    from eyegaze_multimodal_b200.optim import FusedClipAdamW
    from eyegaze_multimodal_b200.parallel import TrialParallel
    from eyegaze_multimodal_b200.precision import set_precision

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback")
    L.load()                                            # fails loudly if libeyegaze_b200.so is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise RuntimeError("--gpus %d needs a torchrun launch (one rank per GPU); WORLD_SIZE is 1" % args.gpus)
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the JSON line must be the only thing on stdout: NCCL_DEBUG=VERSION prints a banner there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    set_precision(args.precision)
    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    train = args.mode == "train"

    model = build_model(wl, dev)
    model.train(train)
    tp = TrialParallel(model, bucket_mb=args.bucket_mb,
                       comm_dtype=torch.bfloat16 if args.grad_comm == "bf16" else torch.float32)
    n_params = sum(p.numel() for p in model.parameters())
    loss_fn = make_loss_fn(wl, tp, train)

    # ---- synthetic inputs (seeded per rank: every rank owns different trials) --------------------------------
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2                                          # distinct pinned host batches cycled by the timed loops
    host = [host_batch(wl, B, g) for _ in range(n_host)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0].values())
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]

    def eager_step(batch):
        if not train:
            with torch.no_grad():
                return loss_fn(model, batch)["loss"]
        tp.zero_grad()
        loss = loss_fn(model, batch)["loss"]
        loss.backward()
        tp.finish()
        return loss

    graphed = None
    if args.graphs:
        if train:
            graphed = GraphedTrainStep(model, loss_fn, resident[0], trial_parallel=tp, warmup=max(2, args.warmup))
        else:
            graphed = GraphedForward(model, lambda m, b: loss_fn(m, b), resident[0], warmup=max(2, args.warmup))

    def run_step(batch):
        """One step on a batch that is already on the device (the graph owns static inputs: one device copy in)."""
        if graphed is None:
            return eager_step(batch)
        out = graphed(batch)
        return out if train else out["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """steps x fn(i), device-timed with CUDA events between two barriers; returns (ms total max over ranks, host ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        host_ms = (time.perf_counter() - w0) * 1e3          # time to ENQUEUE the steps
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), host_ms

    # ---- (1) device-resident throughput ---------------------------------------------------------------------
    for i in range(args.warmup):
        run_step(resident[i % n_host])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.launch_count()
    last = {}

    def resident_step(i):
        last["loss"] = run_step(resident[i % n_host])
    ms_total, host_ms = timed(resident_step, args.steps)
    launches_host = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(last["loss"].item())
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)

    # kernels per step: counted on an eager step (a graph replay issues the same kernels with one host launch)
    n0 = L.launch_count()
    eager_step(resident[0])
    torch.cuda.synchronize()
    kernels_per_step = L.launch_count() - n0
    captured = getattr(graphed, "captured_launches", None) if graphed is not None else None
    gpu_launches = (captured or kernels_per_step) * args.steps if graphed is not None else launches_host

    # ---- (1b) the un-captured step, for comparison: device time and host enqueue time --------------------------
    eager = None
    if graphed is not None:
        for i in range(2):
            eager_step(resident[i % n_host])
        n_eager = min(args.steps, 5)
        ems, ehost = timed(lambda i: eager_step(resident[i % n_host]), n_eager)
        eager = {"ms_per_step": ems / n_eager, "value": B * world * n_eager / (ems * 1e-3), "unit": UNIT,
                 "host_enqueue_ms_per_step": ehost / n_eager, "kernels_per_step": kernels_per_step,
                 "kernels_per_graph_replay": captured,
                 "graph_host_enqueue_ms_per_step": host_ms / args.steps}

    # ---- (2) end to end: pinned host inputs -> H2D -> step -> loss D2H, every step, double-buffered ----------
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        slots = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

        def stage(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])           # slot free once step i-2 has finished with it
                for k, v in host[i % n_host].items():
                    slots[s][k].copy_(v, non_blocking=True)
                ready[s].record(copy_stream)

        def e2e_loop(n):
            for s in range(2):
                consumed[s].record()
            stage(0)
            for i in range(n):
                if i + 1 < n:
                    stage(i + 1)                             # overlaps the previous step's compute
                torch.cuda.current_stream().wait_event(ready[i % 2])
                ls = run_step(slots[i % 2])
                consumed[i % 2].record()
                loss_host.copy_(ls.detach().reshape(1), non_blocking=True)
            torch.cuda.synchronize()                          # the last loss has reached the host
            return float(loss_host[0])

        e2e_loop(2)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        e2e_loop(args.steps)
        t1.record()
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        e2e_ms = max_over_ranks(max(t0.elapsed_time(t1), wall_ms))
        e2e = {"value": B * world * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
               "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms / args.steps,
               "api": "drop-in nn.Module forward + loss + backward%s; pinned host batches, double-buffered H2D on a copy "
                      "stream, loss read back every step" % (" (captured CUDA graph, GraphedTrainStep)" if graphed else "")}

    # ---- (3) roofline of the dominant kernel (tcgen05 GEMM), CUDA events around every launch ---------------
    roofline = None
    if not args.no_roofline and train:
        # every rank runs the profiled steps (they contain the gradient all-reduce); rank 0 reports
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(pk):
            with open(pk) as f:
                peaks = json.load(f)
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PF sustained (of fallback)"
        # the two encoder branches normally overlap on two streams; for the per-kernel roofline they are serialised so
        # that an event pair brackets ONE kernel running alone on the device, as it would under ncu (eager launches:
        # events cannot bracket kernels inside a graph)
        was_concurrent = getattr(model, "concurrent_branches", False)
        if was_concurrent:
            model.concurrent_branches = False
        for i in range(2):
            eager_step(resident[i % n_host])
        torch.cuda.synchronize()
        L.prof_read(0, reset=True)
        L.prof_enable(True)
        n_prof = 2
        for i in range(n_prof):
            eager_step(resident[i % n_host])
        torch.cuda.synchronize()
        L.prof_enable(False)
        if was_concurrent:
            model.concurrent_branches = True
        pr = L.prof_read(0, reset=True)
        if rank == 0 and pr["launches"] > 0 and pr["ms"] > 0:
            ach = pr["flops"] / (pr["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05.mma bf16, TMA operands, TMEM accumulators)",
                        "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "traffic": None,
                        "traffic_note": "not measurable inside this run (dram__bytes needs ncu); the committed ncu --set full "
                                        "captures under profiles/ carry it per launch",
                        "algorithmic_bytes_per_launch": pr["bytes"] / pr["launches"],
                        "peak_source": peak_src, "launches_per_step": pr["launches"] / n_prof,
                        "avg_launch_us": pr["ms"] * 1e3 / pr["launches"],
                        "flops_per_launch": pr["flops"] / pr["launches"],
                        "kernel_ms_per_step": pr["ms"] / n_prof,
                        "share_of_step": pr["ms"] / n_prof / ms_per_step,
                        "note": "launch durations measured with the encoder branches serialised (one kernel on the device "
                                "at a time); share_of_step = their sum / the timed step of `value` (%.2f ms), in which the "
                                "two branches overlap" % ms_per_step,
                        "model_flops_per_step": pr["flops"] / n_prof,
                        "step_tensor_frac": pr["flops"] / n_prof / (ms_per_step * 1e-3) / 1e12 / peak_tf}
    if world > 1:
        dist.barrier()

    # ---- (4) the training-step tail inside the graph: fwd + bwd + global-norm clip + AdamW ---------------------
    train_step = None
    if train and graphed is not None and not args.no_train_step:
        graphed = None                                       # drop the first graph's memory pool
        torch.cuda.empty_cache()
        opt = FusedClipAdamW(model.parameters(), lr=1e-5, weight_decay=0.01, max_grad_norm=1.0, capturable=True)
        full = GraphedTrainStep(model, loss_fn, resident[0], optimizer=opt, trial_parallel=tp, warmup=2)
        for i in range(2):
            full(resident[i % n_host])
        tms, thost = timed(lambda i: full(resident[i % n_host]), args.steps)
        train_step = {"value": B * world * args.steps / (tms * 1e-3), "unit": UNIT, "ms_per_step": tms / args.steps,
                      "host_enqueue_ms_per_step": thost / args.steps,
                      "what": "fwd + 4-term loss + bwd%s + global-norm clip (1.0) + AdamW over %d parameters, one captured "
                              "CUDA graph per step" % (" + gradient all-reduce" if world > 1 else "", n_params)}
        del full
        torch.cuda.empty_cache()

    # ---- (5) baselines on this box: PyTorch-eager on the GPU, the oracle port on the host cores ----------------
    gpu_eager = cpu = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        from oracle import eeg as O
        O.IBS_CHUNK = 32
        gpu_eager = time_gpu_eager(wl, B, dev, steps=3, warmup=1, train=train)
        O.IBS_CHUNK = 4
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tps, s_step, n_threads = time_cpu(wl, args.cpu_sample, steps=2, warmup=1, train=train)
        cpu = {"value": tps, "unit": UNIT, "cores": n_threads, "kind": "port",
               "sample": "%d trials per step x 2 timed steps (1 warm-up) of workload %s, fp32, %s, %.2f s/step"
                         % (args.cpu_sample, args.workload, "fwd+bwd" if train else "fwd only", s_step)}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": make_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
            "train_step": train_step, "eager": eager, "gpu_eager_baseline": gpu_eager,
            "run": {"loss": loss_val, "parameters": n_params, "cuda_graph": bool(args.graphs),
                    "h2d_mb_per_step_per_gpu": h2d_bytes / 1e6},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints while we run (NCCL's version banner, warnings
    # from C code) is diverted to stderr at the file-descriptor level and stdout is restored for the final print.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    sys.stdout = real_stdout_proxy = _Tee(real_stdout)
    try:
        if args.impl == "reference":
            return run_reference(args)
        return run_b200(args)
    finally:
        real_stdout_proxy.flush()


class _Tee:
    """print() goes to the REAL stdout (saved descriptor); C-level writes to fd 1 go to stderr."""

    def __init__(self, f):
        self.f = f

    def write(self, x):
        return self.f.write(x)

    def flush(self):
        self.f.flush()


if __name__ == "__main__":
    sys.exit(main())
