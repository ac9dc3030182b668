"""Trial-wise data parallelism for the drop-in models: one process per GPU, replicated parameters, the batch of
trials split across ranks, and ONE real exchange step per training step -- the gradient all-reduce.

The reference is single-process (no torch.distributed anywhere); trials are independent in
forward and backward (LayerNorm per token, InstanceNorm per trial, no BatchNorm), so the path shards with no
data-path collective.  Gradients are reduced in flat fp32 buckets:

  * ``zero_grad()`` drops the gradients (``.grad = None``), so autograd hands every parameter its gradient tensor
    without an accumulation kernel;
  * a post-accumulate-grad hook counts a bucket's parameters down during backward; when the last one lands, ONE
    multi-tensor copy packs the bucket's gradients into its flat buffer, ``.grad`` is re-pointed at the views, and the
    bucket's all-reduce (NCCL over NVLink / NVSwitch, average) is enqueued asynchronously, so it overlaps the rest
    of the backward pass (buckets are filled in reverse registration order ~ backward execution order);
  * with a single rank nothing is packed or reduced at all;
  * ``finish()`` flushes buckets whose parameters received no gradient this step, makes the compute stream wait
    for the communication and RE-ARMS the buckets, so loops that call ``optimizer.zero_grad()`` / ``model.zero_grad()``
    (as the reference's loops do, train_art.py:216, train_multimodal_fuzzy_fusion.py:434) instead of
    ``TrialParallel.zero_grad()`` keep reducing every step;
  * one backward per ``finish()``: gradient accumulation over several backward passes is not supported and raises
    (a second backward would add local gradients into already reduced buffers);
  * ragged shards (``B % world != 0``): scale each rank's loss by ``loss_weight(local_B, global_B)`` so the averaged
    gradients equal the single-process global-batch mean.

``backend='gloo'`` (CPU tests, world_size 2) takes the same code path with SUM + scale instead of NCCL's AVG.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


class _Bucket:
    __slots__ = ("flat", "params", "views", "pending", "work", "launched", "comm")

    def __init__(self, flat, params, views, comm=None):
        self.flat = flat
        self.params = params
        self.views = views
        self.pending = len(params)
        self.work = None
        self.launched = False
        self.comm = comm            # optional bf16 wire buffer (comm_dtype=torch.bfloat16)


class TrialParallel(nn.Module):
    """Wraps a model; ``forward`` is the model's.  Call ``zero_grad()`` before and ``finish()`` after ``backward()``."""

    def __init__(self, module: nn.Module, bucket_mb: float = 32.0, process_group=None, broadcast: bool = True,
                 comm_dtype: torch.dtype = torch.float32):
        """``comm_dtype=torch.bfloat16`` sends the gradients over the wire in bf16 (half the bytes and half the time the
        NCCL channels share the SMs with backward; the sum over ranks is then rounded to bf16, as with PyTorch DDP's
        ``bf16_compress_hook``); the optimiser still sees fp32 gradients.  Default: fp32, bit-comparable with the
        single-process gradient."""
        super().__init__()
        if comm_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("TrialParallel: comm_dtype must be torch.float32 or torch.bfloat16")
        self.comm_dtype = comm_dtype
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        if broadcast and self.world > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0, group=process_group)
        self.buckets: List[_Bucket] = []
        # True: the hooks only PACK a bucket when its last gradient lands and finish() issues the all-reduces -- the mode
        # of a captured training step (graphs.GraphedTrainStep), where NCCL launches stay outside the CUDA graph
        self.defer_comm = False
        self._build_buckets(int(bucket_mb * (1 << 20)))

    # ------------------------------------------------------------------------------------------------
    def _build_buckets(self, bucket_bytes: int) -> None:
        params = [p for p in self.module.parameters() if p.requires_grad]
        params.reverse()                       # heads first: the order in which backward produces gradients
        groups, cur, cur_bytes = [], [], 0
        for p in params:
            nb = p.numel() * 4
            if cur and (cur_bytes + nb > bucket_bytes or cur[0].device != p.device):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            groups.append(cur)
        for g in groups:
            # 16-byte aligned slices keep the kernels' vectorised access to parameter gradients legal
            offs, n = [], 0
            for p in g:
                if p.dtype != torch.float32:
                    raise TypeError("TrialParallel expects fp32 master parameters")
                offs.append(n)
                n += (p.numel() + 3) // 4 * 4
            flat = torch.zeros(n if self.world > 1 else 1, dtype=torch.float32, device=g[0].device)
            views = [flat[o:o + p.numel()].view(p.shape) for p, o in zip(g, offs)] if self.world > 1 else []
            comm = None
            if self.world > 1 and self.comm_dtype != torch.float32:
                comm = torch.zeros(n, dtype=self.comm_dtype, device=g[0].device)
            b = _Bucket(flat, g, views, comm)
            for p in g:
                p.grad = None
                p.register_post_accumulate_grad_hook(self._make_hook(b))
            self.buckets.append(b)

    def _make_hook(self, b: _Bucket):
        def hook(_p):
            b.pending -= 1
            if b.pending < 0:
                raise RuntimeError("TrialParallel: a parameter received a second gradient before finish() -- one "
                                   "backward pass per finish(); gradient accumulation is not supported")
            if b.pending == 0:
                self._launch(b)
        return hook

    def _launch(self, b: _Bucket) -> None:
        b.launched = True
        if self.world == 1:
            return
        if b.flat.is_cuda:
            # gradients of a branch the model ran on a side stream (MultimodalFusionModel) are produced there
            side = getattr(self.module, "_side_stream", None)
            if side is not None:
                torch.cuda.current_stream(b.flat.device).wait_stream(side)
        with torch.no_grad():
            have = [(v, p.grad) for v, p in zip(b.views, b.params) if p.grad is not None]
            if len(have) != len(b.params):
                b.flat.zero_()                 # parameters the loss did not reach contribute zeros
            have = [(v, g) for v, g in have if g is not v]     # (a replayed graph has already packed into the views)
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            for v, p in zip(b.views, b.params):
                p.grad = v                     # the reduced values are what the optimiser sees
        if not self.defer_comm:
            self._reduce(b)

    def _reduce(self, b: _Bucket) -> None:
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        buf = b.flat
        if b.comm is not None:
            b.comm.copy_(b.flat)               # one cast kernel: the packed fp32 bucket -> its bf16 wire image
            buf = b.comm
        b.work = dist.all_reduce(buf, op=op, group=self.group, async_op=True)

    # ------------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def _rearm(self, b: _Bucket) -> None:
        b.pending = len(b.params)
        b.work = None
        b.launched = False

    def zero_grad(self, set_to_none: bool = True) -> None:   # noqa: D401
        for b in self.buckets:
            for p in b.params:
                p.grad = None
            self._rearm(b)

    def finish(self) -> None:
        """Reduce whatever has not been reduced yet, order the compute stream after all communication and re-arm the
        buckets for the next backward pass."""
        if self.defer_comm and self.world > 1 and torch.cuda.is_current_stream_capturing():
            return                             # capture ends with the packed buckets; the caller reduces after the replay
        for b in self.buckets:
            if not b.launched:
                self._launch(b)
            if self.defer_comm and self.world > 1 and b.work is None:
                self._reduce(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                if b.comm is not None:
                    b.flat.copy_(b.comm)       # back to the fp32 views the optimiser reads
                if not self._avg:
                    b.flat.div_(self.world)
            if self.world > 1:
                for v, p in zip(b.views, b.params):
                    if p.grad is not v:
                        raise RuntimeError("TrialParallel.finish(): a gradient was replaced after its bucket had been "
                                           "reduced (second backward pass before finish()?)")
            self._rearm(b)

    # -- batch-level auxiliary losses (dual_eeg_transformer.py:1255-1371) look ACROSS the batch: gather the shards first
    def compute_symmetry_loss(self, cls1, cls2):
        return self._inner().compute_symmetry_loss(cls1, cls2)        # a mean over trials: per shard is exact for even shards

    def compute_ibs_alignment_loss(self, ibs_token, cls1, cls2, temperature: float = 0.07):
        g = self.group
        return self._inner().compute_ibs_alignment_loss(gather_trials(ibs_token, g), gather_trials(cls1, g),
                                                        gather_trials(cls2, g), temperature)

    def compute_ibs_contrastive_loss(self, ibs_tokens, labels, temperature: float = 0.07):
        g = self.group
        return self._inner().compute_ibs_contrastive_loss(gather_trials(ibs_tokens, g), gather_trials(labels, g),
                                                          temperature)

    def _inner(self):
        m = self.module
        return getattr(m, "eeg_encoder", m)

    def grad_bytes(self) -> int:
        """Bytes one rank hands to the all-reduce per step."""
        esz = 2 if self.comm_dtype == torch.bfloat16 else 4
        return sum(sum(p.numel() for p in b.params) * esz for b in self.buckets)

    def flat_grads(self) -> Iterable[torch.Tensor]:
        return [b.flat for b in self.buckets]


class _GatherTrials(torch.autograd.Function):
    """All-gather of per-trial rows along dim 0 (equal shards).  Every rank then evaluates the SAME global loss, so the
    gradient of that loss with respect to this rank's rows is the local slice of the incoming gradient, times `world`
    because the gradient all-reduce that follows AVERAGES over ranks."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        ctx.meta = (rank, world, x.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        rank, world, n = ctx.meta
        return g[rank * n:(rank + 1) * n] * float(world), None


def gather_trials(x: torch.Tensor, group=None) -> torch.Tensor:
    """(B_local, ...) -> (B_global, ...) across the ranks of `group` (identity for a single process)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    if x.is_floating_point() and x.requires_grad:
        return _GatherTrials.apply(x, group)
    out = torch.empty((dist.get_world_size(group) * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def loss_weight(local_trials: int, global_trials: int, world: Optional[int] = None) -> float:
    """Factor for a rank's batch-MEAN loss so that the AVG all-reduce of the gradients equals the gradient of the
    global-batch mean when shards are ragged: (local_B / global_B) * world  (= 1 for even shards)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return float(local_trials) * world / float(global_trials)


def shard_trials(n_trials: int, rank: Optional[int] = None, world: Optional[int] = None) -> range:
    """Contiguous shard of trial indices owned by ``rank`` (ragged tails go to the low ranks)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    base, rem = divmod(n_trials, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))
