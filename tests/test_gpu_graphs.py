"""CUDA-graph capture of the training step (graphs.GraphedTrainStep / GraphedForward): a replay must compute what the
eager step computes, draw fresh dropout masks on every replay, and carry the captured optimizer tail.
Run on the B200 box:  pytest -m gpu"""
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200 import _lib as L
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.early_fusion_vit import EarlyFusionViT
    from eyegaze_multimodal_b200.fuzzy_gating_fusion import FuzzyGatingFusion
    from eyegaze_multimodal_b200.graphs import GraphedForward, GraphedTrainStep
    from eyegaze_multimodal_b200.multimodal import MultimodalFusionModel, multimodal_loss
    from eyegaze_multimodal_b200.optim import DeviceLRSchedule, FusedClipAdamW
    from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import eeg_pair_batch, gaze_pair_batch
from oracle import eeg as O
from oracle import fuzzy as FZ
from oracle import vit as V

DEV = "cuda:0"
NAME = "vit_tiny_patch16_224"
B, C, T = 4, 8, 256


def _model(seed=0):
    warnings.simplefilter("ignore")
    cfg = O.EEGConfig(in_channels=C, d_model=64, num_layers=2, num_heads=4, d_ff=128, max_len=T // 2)
    eeg = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    eeg.load_state_dict(O.init_state_dict(cfg, 7 + seed), strict=True)
    gaze = EarlyFusionViT(NAME, pretrained=False, fusion_mode="concat")
    gaze.load_state_dict(V.init_vit_state_dict(NAME, 6, 3, "backbone.", seed=8 + seed), strict=True)
    return MultimodalFusionModel(gaze, eeg, FuzzyGatingFusion(3, "full")).to(DEV)


def _batch(seed=9):
    e1, e2 = eeg_pair_batch(B, C, T, seed=seed, coupled=True)
    a, b = gaze_pair_batch(B, seed=seed + 1)
    return {"img1": a.to(DEV), "img2": b.to(DEV), "eeg1": e1.to(DEV), "eeg2": e2.to(DEV),
            "labels": (torch.arange(B) % 3).to(DEV)}


def _loss_fn(model, batch):
    out = model(batch["img1"], batch["img2"], batch["eeg1"], batch["eeg2"], batch["labels"])
    return {"loss": multimodal_loss(model, out, batch["labels"]), "fused_logits": out["fused_logits"]}


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_replay_equals_eager_step(cuda_device, mode):
    model = _model().eval()                          # no dropout: replays are deterministic
    batch = _batch()
    with precision(mode):
        model.zero_grad(set_to_none=True)
        want = _loss_fn(model, batch)
        want["loss"].backward()
        torch.cuda.synchronize()
        ref_loss = want["loss"].detach().clone()
        ref_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        # nothing may keep the eager autograd graph alive: its AccumulateGrad nodes were created on the default (legacy)
        # stream and would run there again during capture, which CUDA refuses
        del want
        model.zero_grad(set_to_none=True)
        step = GraphedTrainStep(model, _loss_fn, batch)
        n0 = L.launch_count()
        for _ in range(2):
            loss = step.replay()
        torch.cuda.synchronize()
        assert L.launch_count() == n0                 # a replay issues no launch from the host side of the library
    assert torch.allclose(loss, ref_loss, rtol=1e-6, atol=1e-7), (float(loss), float(ref_loss))
    got = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(ref_grads)
    for n, g in ref_grads.items():
        tol = 1e-5 * g.abs().max().item() + 1e-9 if mode == "fp32" else 2e-2 * g.abs().max().item() + 1e-7
        assert (got[n] - g).abs().max().item() <= tol, n      # (split-K accumulates with atomics: not bit-exact)
    # a different batch through the same graph
    other = _batch(seed=30)
    with precision(mode):
        model.zero_grad(set_to_none=True)
        want2 = _loss_fn(model, other)["loss"].detach().clone()
    loss2 = step(other)
    torch.cuda.synchronize()
    assert torch.allclose(loss2, want2, rtol=1e-5, atol=1e-6) and not torch.allclose(loss2, ref_loss, rtol=1e-4)


def test_replays_draw_fresh_dropout_masks(cuda_device):
    model = _model().train()
    batch = _batch()
    with precision("bf16"):
        step = GraphedTrainStep(model, _loss_fn, batch)
        losses = []
        for _ in range(4):
            losses.append(float(step.replay()))
    assert len(set(losses)) == 4, losses              # the device seed epoch advances inside the graph
    assert all(abs(x - losses[0]) < 0.5 for x in losses)


def test_captured_optimizer_and_schedule_match_eager_training(cuda_device):
    """N replays of [fwd, bwd, clip + AdamW, LR schedule] == N eager steps (eval mode: no dropout randomness)."""
    kw = dict(lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    batch = _batch()
    ref = _model().eval()
    ropt = FusedClipAdamW(ref.parameters(), **kw)
    rsch = torch.optim.lr_scheduler.LambdaLR(ropt, lambda s: FZ_factor(s))
    n_steps = 5
    with precision("fp32"):
        for _ in range(n_steps):
            ropt.zero_grad(set_to_none=True)
            _loss_fn(ref, batch)["loss"].backward()
            ropt.step()
            rsch.step()
    torch.cuda.synchronize()
    model = _model().eval()
    opt = FusedClipAdamW(model.parameters(), capturable=True, **kw)
    sch = DeviceLRSchedule(opt, "warmup_cosine", warmup_steps=2, total_steps=12)
    with precision("fp32"):
        # GraphedTrainStep's warm-up trains too: restore the initial state before the counted replays
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        step = GraphedTrainStep(model, _loss_fn, batch, optimizer=opt, schedule=sch, warmup=1)
        model.load_state_dict(sd0)
        opt.state_dict()                              # (sync host view)
        for st in opt.state.values():
            st["exp_avg"].zero_()
            st["exp_avg_sq"].zero_()
        opt.device_state()[:3] = 0.0
        sch.load_state_dict({"kind": "warmup_cosine", "p0": 2.0, "p1": 12.0, "last_epoch": 0.0, "opt_step": 0.0})
        from eyegaze_multimodal_b200 import ops
        ops.bump_param_epoch()
        for _ in range(n_steps):
            step.replay()
    torch.cuda.synchronize()
    for (n, p), (_, r) in zip(model.named_parameters(), ref.named_parameters()):
        # (the reductions behind several gradients use fp32 atomics, whose order differs between launches; Adam turns a
        #  gradient that is analytically zero -- the key bias of every attention layer -- into O(lr) steps whose direction
        #  follows that rounding noise.  1e-4 is a tenth of ONE step at lr = 1e-3; measured differences are <= 2.1e-5.)
        tol = 1e-4 * max(1.0, r.detach().abs().max().item())
        assert (p.detach() - r.detach()).abs().max().item() <= tol, n
    assert abs(opt.device_step().item() - n_steps) < 1e-6


def FZ_factor(s, warm=2, total=12):
    import math
    if s < warm:
        return float(s) / float(max(1, warm))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(s - warm) / float(max(1, total - warm)))))


def test_graphed_forward_equals_eager_inference(cuda_device):
    model = _model().eval()
    batch = _batch()

    def fwd(m, b):
        return m(b["img1"], b["img2"], b["eeg1"], b["eeg2"])["fused_logits"]
    with precision("bf16"), torch.no_grad():
        want = fwd(model, batch).clone()
        g = GraphedForward(model, fwd, batch)
        got = g()["out"]
        torch.cuda.synchronize()
        assert torch.equal(got, want)
        other = _batch(seed=40)
        want2 = fwd(model, other).clone()
        got2 = g(other)["out"]
        torch.cuda.synchronize()
    assert torch.equal(got2, want2) and not torch.equal(want, want2)
