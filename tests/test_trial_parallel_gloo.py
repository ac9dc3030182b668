"""world_size-2 gloo test (CPU) of the trial-parallel gradient exchange: sharding trials over ranks and averaging the
bucketed gradients must reproduce the single-process gradient of the full batch; parameters used twice, unused
parameters and ragged shards are covered."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from eyegaze_multimodal_b200.parallel import TrialParallel, loss_weight, shard_trials


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = nn.Linear(16, 32)         # applied to both "players": one parameter, two uses per step
        self.head = nn.Sequential(nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 3))
        self.unused = nn.Linear(4, 4)        # never reached by the loss

    def forward(self, a, b):
        return self.head(torch.cat([torch.tanh(self.enc(a)), torch.tanh(self.enc(b))], -1))


def _data(n=10):
    g = torch.Generator().manual_seed(3)
    return torch.randn(n, 16, generator=g), torch.randn(n, 16, generator=g), torch.randint(0, 3, (n,), generator=g)


def _worker(rank, world, port, out, comm_dtype=torch.float32):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)            # different init per rank: the wrapper must broadcast rank 0's weights
    model = Toy()
    tp = TrialParallel(model, bucket_mb=0.004, comm_dtype=comm_dtype)   # tiny buckets: several all-reduces per step
    assert len(tp.buckets) > 2
    a, b, y = _data()
    idx = list(shard_trials(len(y), rank, world))
    for _ in range(2):                       # second step: buckets re-armed by zero_grad
        tp.zero_grad()
        # sum-reduced per-shard loss / global count == mean over the full batch after averaging across ranks
        loss = nn.functional.cross_entropy(tp(a[idx], b[idx]), y[idx], reduction="sum") * world / len(y)
        loss.backward()
        tp.finish()
    out[rank] = {n: p.grad.clone() for n, p in model.named_parameters()} | {"w0": model.enc.weight.detach().clone()}
    dist.destroy_process_group()


def test_gradients_match_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    torch.manual_seed(100)
    ref = Toy()
    a, b, y = _data()
    nn.functional.cross_entropy(ref(a, b), y).backward()
    assert torch.equal(out[0]["w0"], out[1]["w0"]) and torch.equal(out[0]["w0"], ref.enc.weight.detach())
    for n, p in ref.named_parameters():
        want = p.grad if p.grad is not None else torch.zeros_like(p)
        for r in (0, 1):
            assert torch.allclose(out[r][n], want, atol=1e-6), (n, r)


def test_bf16_wire_gradients_close_to_single_process():
    """comm_dtype=bfloat16: the buckets cross the wire in bf16, the optimiser still sees fp32 tensors; the result is the
    single-process gradient up to bf16 rounding of each rank's contribution and of the sum."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out, torch.bfloat16), nprocs=2, join=True)
    torch.manual_seed(100)
    ref = Toy()
    a, b, y = _data()
    nn.functional.cross_entropy(ref(a, b), y).backward()
    for n, p in ref.named_parameters():
        want = p.grad if p.grad is not None else torch.zeros_like(p)
        for r in (0, 1):
            assert out[r][n].dtype == torch.float32
            assert torch.allclose(out[r][n], want, atol=2e-2 * float(want.abs().max()) + 1e-6), (n, r)
        assert torch.equal(out[0][n], out[1][n])        # every rank ends with the same reduced gradient


def test_shard_trials_ragged():
    assert [list(shard_trials(10, r, 4)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert list(shard_trials(3, 3, 4)) == []
    assert sum(len(shard_trials(4096, r, 8)) for r in range(8)) == 4096


def _worker_reference_loop(rank, world, port, out):
    """The reference's loops call optimizer.zero_grad() / model.zero_grad(), never TrialParallel.zero_grad()
    (train_art.py:216): finish() must re-arm the buckets by itself.  Shards are RAGGED (7 trials over 2 ranks) and each
    rank uses its batch-MEAN loss scaled by loss_weight(), so the averaged gradients equal the global-batch mean."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100)
    model = Toy()
    tp = TrialParallel(model, bucket_mb=0.004)
    a, b, y = _data(7)
    idx = list(shard_trials(len(y), rank, world))
    seen = []
    for step in range(3):
        model.zero_grad(set_to_none=True)                # NOT tp.zero_grad()
        loss = nn.functional.cross_entropy(tp(a[idx], b[idx]), y[idx]) * loss_weight(len(idx), len(y), world)
        loss.backward()
        tp.finish()
        seen.append({n: p.grad.clone() for n, p in model.named_parameters()})
    # a second backward before finish() must fail loudly instead of silently skipping the reduction
    model.zero_grad(set_to_none=True)
    nn.functional.cross_entropy(tp(a[idx], b[idx]), y[idx]).backward()
    try:
        nn.functional.cross_entropy(tp(a[idx], b[idx]), y[idx]).backward()
        raised = False
    except RuntimeError as e:
        raised = "gradient accumulation is not supported" in str(e)
    out[rank] = {"steps": seen, "raised": raised}
    dist.destroy_process_group()


def test_reference_style_loop_rearms_and_ragged_shards():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_reference_loop, args=(2, port, out), nprocs=2, join=True)
    torch.manual_seed(100)
    ref = Toy()
    a, b, y = _data(7)
    nn.functional.cross_entropy(ref(a, b), y).backward()
    for r in (0, 1):
        assert out[r]["raised"]
        for step in range(3):                            # every step reduced: all three equal the full-batch gradient
            for n, p in ref.named_parameters():
                want = p.grad if p.grad is not None else torch.zeros_like(p)
                assert torch.allclose(out[r]["steps"][step][n], want, atol=1e-6), (n, r, step)


def test_loss_weight():
    assert loss_weight(4, 8, 2) == 1.0
    assert abs(loss_weight(4, 7, 2) + loss_weight(3, 7, 2) - 2.0) < 1e-12


def _batch_loss(tok, y):
    """A loss that looks ACROSS the batch (stand-in for the contrastive aux losses, dual_eeg_transformer.py:1305-1371)."""
    t = nn.functional.normalize(tok, dim=1)
    sim = t @ t.t() / 0.5
    same = (y[:, None] == y[None, :]).float()
    return (torch.logsumexp(sim, 1) - (sim * same).sum(1) / same.sum(1)).mean()


def _worker_gather(rank, world, port, out):
    from eyegaze_multimodal_b200.parallel import gather_trials
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100)
    model = Toy()
    tp = TrialParallel(model, bucket_mb=0.004)
    a, b, y = _data(8)
    idx = list(shard_trials(len(y), rank, world))
    tp.zero_grad()
    tok = tp(a[idx], b[idx])                                 # (B_local, 3) "tokens"
    loss = _batch_loss(gather_trials(tok), gather_trials(y[idx]))
    loss.backward()
    tp.finish()
    out[rank] = {"loss": loss.detach().clone(), "grads": {n: p.grad.clone() for n, p in model.named_parameters()}}
    dist.destroy_process_group()


def test_gather_trials_makes_batch_level_losses_exact_under_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_gather, args=(2, port, out), nprocs=2, join=True)
    torch.manual_seed(100)
    ref = Toy()
    a, b, y = _data(8)
    loss = _batch_loss(ref(a, b), y)
    loss.backward()
    for r in (0, 1):
        assert torch.allclose(out[r]["loss"], loss.detach(), atol=1e-6)
        for n, p in ref.named_parameters():
            want = p.grad if p.grad is not None else torch.zeros_like(p)
            assert torch.allclose(out[r]["grads"][n], want, atol=1e-6), (n, r)
