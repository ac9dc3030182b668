"""Fused global-norm clipping + AdamW against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW
run on the CPU in fp64-free plain fp32 (the reference's training loops call exactly these two,
train_art.py:221-229).  Run on the B200 box:  pytest -m gpu"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200.optim import FusedClipAdamW

DEV = "cuda:0"
SHAPES = [(1,), (3,), (7, 5), (16384,), (16385,), (257, 129), (768, 768), (40000,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g) * 0.3 for s in SHAPES]


@pytest.mark.parametrize("max_norm", [None, 1.0, 1e4])
def test_matches_torch_clip_and_adamw(cuda_device, max_norm):
    ref = [torch.nn.Parameter(t.clone()) for t in _params(0)]
    mine = [torch.nn.Parameter(t.clone().to(DEV)) for t in _params(0)]
    kw = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    o_ref = torch.optim.AdamW(ref, **kw)
    o_mine = FusedClipAdamW(mine, max_grad_norm=max_norm, **kw)
    for step in range(4):
        grads = [t * (5.0 if step == 1 else 0.2) for t in _params(10 + step)]
        for i, (r, m, g) in enumerate(zip(ref, mine, grads)):
            skip = (step == 2 and i == 3)                 # a parameter without gradient is left alone, like torch
            r.grad = None if skip else g.clone()
            m.grad = None if skip else g.clone().to(DEV)
        total = None
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        o_ref.step()
        o_mine.step()
        if max_norm is not None:
            assert abs(o_mine.grad_norm().item() - total.item()) <= 1e-5 * total.item()
        for r, m, s in zip(ref, mine, SHAPES):
            err = (m.detach().cpu() - r.detach()).abs().max().item()
            assert err <= 2e-6 * max(1.0, r.detach().abs().max().item()), (step, s, err)
    # AdamW's state_dict layout: moments agree and load into torch's optimizer
    sd = o_mine.state_dict()
    for i, r in enumerate(ref):
        st = sd["state"][i]
        assert set(st) == {"step", "exp_avg", "exp_avg_sq"}
        assert (st["exp_avg"].cpu() - o_ref.state[r]["exp_avg"]).abs().max() <= 1e-6
        assert (st["exp_avg_sq"].cpu() - o_ref.state[r]["exp_avg_sq"]).abs().max() <= 1e-6


def test_state_dict_round_trip_and_weight_cache(cuda_device):
    """Resuming from a state_dict continues the same trajectory, and the bf16 weight copies of the drop-in modules are
    refreshed after a step (the kernels write parameters behind autograd's version counters)."""
    from eyegaze_multimodal_b200 import ops
    w = torch.nn.Parameter(torch.randn(64, 32, device=DEV))
    a = FusedClipAdamW([w], lr=1e-2, max_grad_norm=1.0)
    before = ops.weight_plain(w, ops.BF16).clone()
    w.grad = torch.randn_like(w)
    a.step()
    after = ops.weight_plain(w, ops.BF16)
    assert not torch.equal(before, after) and torch.equal(after, w.detach().bfloat16())
    sd = a.state_dict()
    w2 = torch.nn.Parameter(w.detach().clone())
    b = FusedClipAdamW([w2], lr=1e-2, max_grad_norm=1.0)
    b.load_state_dict(sd)
    g = torch.randn_like(w)
    w.grad, w2.grad = g.clone(), g.clone()
    a.step()
    b.step()
    assert torch.allclose(w, w2, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------------
# rest of SURVEY 8f rank 1: GradScaler protocol, non-finite skip, device LR schedules with per-group rates, metric sums
# ---------------------------------------------------------------------------------------------------------------------
def _adamw_reference(ps, grads_per_step, lrs_per_step, kw, max_norm=None):
    """torch.optim.AdamW (+ clip_grad_norm_) on the CPU with an explicit learning rate per step and group."""
    ref = [[torch.nn.Parameter(t.clone()) for t in grp] for grp in ps]
    opt = torch.optim.AdamW([{"params": g, "lr": 1.0} for g in ref], **kw)
    for grads, lrs in zip(grads_per_step, lrs_per_step):
        for grp, gg, pg, lr in zip(ref, grads, opt.param_groups, lrs):
            pg["lr"] = lr
            for p, g in zip(grp, gg):
                p.grad = g.clone()
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_([p for grp in ref for p in grp], max_norm)
        opt.step()
    return ref


def test_gradscaler_protocol_unscale_and_skip_on_device(cuda_device):
    """train_multimodal_fuzzy_fusion.py:462-472: scaler.scale(loss).backward(); [unscale_; clip]; scaler.step(opt);
    scaler.update().  With `_step_supports_amp_scaling` the scale / found_inf reach the kernels as device tensors."""
    kw = dict(betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    shapes = [(33,), (257, 129), (16384,)]
    g = torch.Generator().manual_seed(0)
    p0 = [torch.randn(*s, generator=g) * 0.3 for s in shapes]
    grads = [[torch.randn(*s, generator=g) for s in shapes] for _ in range(3)]
    mine = [torch.nn.Parameter(t.clone().to(DEV)) for t in p0]
    opt = FusedClipAdamW(mine, lr=2e-3, max_grad_norm=1.0, **kw)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=1000)
    scaler.scale(torch.zeros((), device=DEV))                     # lazily creates the scale tensor
    snapshots = []
    for step, gs in enumerate(grads):
        for p, gr in zip(mine, gs):
            p.grad = (gr * scaler.get_scale()).to(DEV)             # what scaler.scale(loss).backward() leaves behind
        if step == 1:
            mine[1].grad[3, 5] = float("inf")                      # an overflowing step: must be skipped entirely
        scaler.step(opt)
        scaler.update()
        snapshots.append([p.detach().cpu().clone() for p in mine])
    assert scaler.get_scale() == 512.0                             # backed off once
    assert all(torch.equal(a, b) for a, b in zip(snapshots[0], snapshots[1]))    # step 1 changed nothing
    ref = _adamw_reference([p0], [[grads[0]], [grads[2]]], [[2e-3], [2e-3]], kw, max_norm=1.0)[0]
    for r, m in zip(ref, mine):
        assert (m.detach().cpu() - r.detach()).abs().max() <= 3e-6
    # reference-style order: unscale_ + torch's clip first, then step (stage UNSCALED: grad_scale is None)
    mine2 = [torch.nn.Parameter(t.clone().to(DEV)) for t in p0]
    opt2 = FusedClipAdamW(mine2, lr=2e-3, **kw)
    sc2 = torch.amp.GradScaler("cuda", init_scale=256.0)
    sc2.scale(torch.zeros((), device=DEV))
    for p, gr in zip(mine2, grads[0]):
        p.grad = (gr * 256.0).to(DEV)
    sc2.unscale_(opt2)
    torch.nn.utils.clip_grad_norm_(mine2, 1.0)
    sc2.step(opt2)
    sc2.update()
    ref2 = _adamw_reference([p0], [[grads[0]]], [[2e-3]], kw, max_norm=1.0)[0]
    for r, m in zip(ref2, mine2):
        assert (m.detach().cpu() - r.detach()).abs().max() <= 3e-6


def test_skip_nonfinite_uses_our_own_norm(cuda_device):
    w = torch.nn.Parameter(torch.randn(1000, device=DEV))
    before = w.detach().clone()
    opt = FusedClipAdamW([w], lr=1e-2, max_grad_norm=1.0, skip_nonfinite=True)
    w.grad = torch.randn(1000, device=DEV)
    w.grad[7] = float("nan")
    opt.step()
    assert torch.equal(w.detach(), before) and not torch.isfinite(opt.grad_norm()).item()
    w.grad = torch.randn(1000, device=DEV)
    opt.step()
    assert not torch.equal(w.detach(), before) and torch.isfinite(w).all()


@pytest.mark.parametrize("kind", ["warmup_cosine", "cosine"])
def test_device_lr_schedule_two_param_groups_matches_torch(cuda_device, kind):
    """Two learning-rate groups (encoder 1e-5-style / fusion 1e-4-style, train_multimodal_fuzzy_fusion.py:727-736) under
    the warm-up + cosine LambdaLR (:197-214, stepped per optimizer step) and CosineAnnealingLR (train_art.py:401-409)."""
    from eyegaze_multimodal_b200.optim import DeviceLRSchedule
    kw = dict(betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    g = torch.Generator().manual_seed(1)
    groups = [[torch.randn(300, generator=g), torch.randn(17, 9, generator=g)], [torch.randn(13, generator=g)]]
    base = [3e-3, 3e-2]
    n_steps = 9
    grads = [[[torch.randn_like(t) for t in grp] for grp in groups] for _ in range(n_steps)]
    # torch: the scheduler the reference builds, on a CPU AdamW
    ref = [[torch.nn.Parameter(t.clone()) for t in grp] for grp in groups]
    topt = torch.optim.AdamW([{"params": r, "lr": lr} for r, lr in zip(ref, base)], **kw)
    if kind == "warmup_cosine":
        warm, total = 3, 8

        def lr_lambda(s):
            if s < warm:
                return float(s) / float(max(1, warm))
            return max(0.0, 0.5 * (1.0 + np.cos(np.pi * float(s - warm) / float(max(1, total - warm)))))
        tsch = torch.optim.lr_scheduler.LambdaLR(topt, lr_lambda)
    else:
        tsch = torch.optim.lr_scheduler.CosineAnnealingLR(topt, T_max=6)
    mine = [[torch.nn.Parameter(t.clone().to(DEV)) for t in grp] for grp in groups]
    opt = FusedClipAdamW([{"params": m, "lr": lr} for m, lr in zip(mine, base)], max_grad_norm=0.5, capturable=True, **kw)
    if kind == "warmup_cosine":
        sch = DeviceLRSchedule(opt, "warmup_cosine", warmup_steps=3, total_steps=8)
    else:
        sch = DeviceLRSchedule(opt, "cosine", T_max=6)
    for step in range(n_steps):
        for grp, mg, gg in zip(ref, mine, grads[step]):
            for p, m, gr in zip(grp, mg, gg):
                p.grad = gr.clone()
                if m.grad is None:
                    m.grad = gr.clone().to(DEV)
                else:
                    m.grad.copy_(gr)                              # capturable: gradients keep their addresses
        want_lr = [pg["lr"] for pg in topt.param_groups]
        got_lr = sch.get_last_lr()
        assert np.allclose(got_lr, want_lr, rtol=2e-6, atol=1e-12), (step, got_lr, want_lr)
        torch.nn.utils.clip_grad_norm_([p for grp in ref for p in grp], 0.5)
        topt.step()
        tsch.step()
        opt.step()
        sch.step()
        for grp, mg in zip(ref, mine):
            for p, m in zip(grp, mg):
                assert (m.detach().cpu() - p.detach()).abs().max() <= 3e-6 * max(1.0, p.detach().abs().max().item()), step
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == n_steps


def test_metric_accumulator_single_host_read(cuda_device):
    from eyegaze_multimodal_b200.optim import MetricAccumulator
    acc = MetricAccumulator(["loss", "loss_ce", "loss_ibs_cls"], DEV)
    g = torch.Generator().manual_seed(2)
    tot = {"loss": 0.0, "loss_ce": 0.0, "loss_ibs_cls": 0.0}
    hits = n = 0
    preds_dev = torch.empty(37, dtype=torch.int64, device=DEV)
    for step in range(5):
        vals = {k: torch.rand((), generator=g) for k in tot}
        for k in tot:
            tot[k] += float(vals[k])
        logits = torch.randn(37, 3, generator=g)
        labels = torch.randint(0, 3, (37,), generator=g)
        acc.add(loss=vals["loss"].to(DEV), loss_ce=vals["loss_ce"].to(DEV), loss_ibs_cls=vals["loss_ibs_cls"].to(DEV))
        acc.add_predictions(logits.to(DEV), labels.to(DEV), preds_dev)
        assert torch.equal(preds_dev.cpu(), logits.argmax(-1))
        hits += int((logits.argmax(-1) == labels).sum())
        n += 37
    r = acc.result()
    assert r["batches"] == 5 and abs(r["accuracy"] - hits / n) < 1e-6
    for k in tot:
        assert abs(r[k] - tot[k] / 5) < 1e-6
