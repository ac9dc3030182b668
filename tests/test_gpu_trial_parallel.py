"""Two-GPU NCCL test of the trial-parallel gradient exchange on the real model: trials split over two ranks, bucketed
gradients averaged over NVLink, must reproduce the single-GPU gradient of the full batch (fp32 parity mode, eval so
no dropout draw differs).  Skipped on a one-GPU box; run with  gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_trial_parallel.py"""
import os
import socket
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B, C, T, NAME = 4, 8, 256, "vit_tiny_patch16_224"


def _build(dev):
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    from eyegaze_multimodal_b200.multimodal import MultimodalFusionModel
    from oracle import eeg as O
    from oracle import vit as V
    cfg = O.EEGConfig(in_channels=C, d_model=64, num_layers=2, num_heads=4, d_ff=128, max_len=T // 2)
    eeg = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    eeg.load_state_dict(O.init_state_dict(cfg, 7), strict=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gaze = EarlyFusionViT(NAME, pretrained=False, fusion_mode="concat")
    gaze.load_state_dict(V.init_vit_state_dict(NAME, 6, 3, "backbone.", seed=8), strict=True)
    return MultimodalFusionModel(gaze, eeg, FuzzyGatingFusion(3, "full")).to(dev).eval()


def _batch():
    from eyegaze_multimodal_b200.synth import eeg_pair_batch, gaze_pair_batch
    e1, e2 = eeg_pair_batch(B, C, T, seed=9, coupled=True)
    a, b = gaze_pair_batch(B, seed=10)
    return a, b, e1, e2, torch.tensor([0, 1, 2, 1])


def _step(model, tp, dev, idx, precision_mode):
    from eyegaze_multimodal_b200.multimodal import multimodal_loss
    from eyegaze_multimodal_b200.precision import precision
    a, b, e1, e2, y = (t[idx].to(dev) for t in _batch())
    with precision(precision_mode):
        if tp is not None:
            tp.zero_grad()
        out = (tp or model)(a, b, e1, e2, y)
        multimodal_loss(model, out, y).backward()
        if tp is not None:
            tp.finish()
    torch.cuda.synchronize(dev)
    return {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}


def _worker(rank, world, port, out):
    from eyegaze_multimodal_b200.parallel import TrialParallel, shard_trials
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model = _build(dev)
    tp = TrialParallel(model, bucket_mb=1.0)
    assert len(tp.buckets) > 2
    idx = list(shard_trials(B, rank, world))
    for _ in range(2):                                   # the second step re-arms the buckets
        grads = _step(model, tp, dev, idx, "fp32")
    out[rank] = grads
    dist.destroy_process_group()


def test_two_gpu_gradients_match_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    ref = _step(_build(dev), None, dev, list(range(B)), "fp32")
    assert set(out[0]) == set(out[1]) and set(ref) <= set(out[0])
    for n, want in ref.items():
        tol = 2e-4 * want.abs().max().item() + 1e-7
        for r in (0, 1):
            err = (out[r][n] - want).abs().max().item()
            assert err <= tol, (n, r, err, tol)
        assert torch.equal(out[0][n], out[1][n]), n       # both ranks hold the same reduced gradient
