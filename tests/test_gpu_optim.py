"""Fused global-norm clipping + AdamW against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW
run on the CPU in fp64-free plain fp32 (the reference's training loops call exactly these two,
train_art.py:221-229).  Run on the B200 box:  pytest -m gpu"""
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
