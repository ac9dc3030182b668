"""CUDA-graph capture of one training step of the hot path: zero-grad -> forward -> loss -> backward [-> gradient packing]
[-> clip + AdamW], replayed with ONE launch from the host.

Why: a cfg2 step is ~2500 kernel launches issued from Python (~23 ms of one host core against ~52 ms of GPU work, 8 ms of
pure launch latency for the launch-bound EEG-only cfg1).  The reference's loops (train_art.py:160-229,
train_multimodal_fuzzy_fusion.py:425-517) are eager PyTorch; this is the B200-side replacement for that inner loop body
-- "CUDA streams and graphs instead of a tracing compiler".  What had to become capture-safe:

  * dropout seeds are launch arguments (frozen by capture): every mask-drawing kernel mixes a device-resident epoch word
    into its seed and the graph advances that word first thing (ops.advance_seed_epoch) -- forward and backward of one
    replay agree, successive replays draw fresh masks;
  * the small-accumulator arena (ops.small_zeros) is reset at capture boundaries, so every accumulator block used by the
    graph is zeroed by a memset node INSIDE the graph;
  * derived parameter copies (bf16 casts, packed QKV, conv re-layouts) are invalidated before capture, so the graph
    re-derives them from the fp32 master parameters on every replay (they change under the captured optimizer step);
  * the optimizer runs in ``capturable`` mode: step count and learning rates live in device memory.

Multi-GPU (parallel.TrialParallel): the graph ends with the bucket packing; the gradient all-reduces are issued right
after the replay (NCCL launches stay outside the graph), then the optimizer graph runs.
"""
from typing import Callable, Dict, Optional

import torch

from . import ops


class GraphedTrainStep:
    """Captures ``loss = loss_fn(model, batch)``, ``loss.backward()`` and (optionally) ``optimizer.step()``.

    ``example_batch``: dict of CUDA tensors with the shapes / dtypes of every later batch; the step owns static copies
    (``self.batch``) that ``load()`` refills.  ``loss_fn(model, batch) -> scalar tensor`` (or a dict with key 'loss';
    every tensor in the dict becomes a static output readable after ``replay()``).

    No autograd graph of an earlier EAGER step on the default stream may still be alive (e.g. a kept ``loss`` tensor): its
    AccumulateGrad nodes would run on the legacy default stream during capture, which CUDA rejects
    (cudaErrorStreamCaptureImplicit).  Drop such tensors first.
    """

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, example_batch: Dict[str, torch.Tensor],
                 optimizer=None, schedule=None, trial_parallel=None, warmup: int = 3):
        dev = next(iter(example_batch.values())).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors; there is no CPU path")
        self.model, self.loss_fn, self.optimizer, self.schedule, self.tp = model, loss_fn, optimizer, schedule, trial_parallel
        self.batch = {k: v.clone() for k, v in example_batch.items()}
        self.outputs: Dict[str, torch.Tensor] = {}
        self.graph = torch.cuda.CUDAGraph()
        self.opt_graph: Optional[torch.cuda.CUDAGraph] = None
        self.replays = 0
        if optimizer is not None and not getattr(optimizer, "capturable", False):
            raise ValueError("the captured optimizer must be a FusedClipAdamW(capturable=True)")
        ops.enable_seed_epoch()
        params = [p for p in model.parameters() if p.requires_grad]
        multi = trial_parallel is not None and trial_parallel.world > 1

        # ---- eager warm-up on a side stream (allocator / cuBLAS-free; builds twiddles, tensor maps, optimizer plans) ----
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._zero(params)
                self._fwd_bwd()
                if trial_parallel is not None:
                    trial_parallel.finish()
                if optimizer is not None:
                    optimizer.step()
                if schedule is not None:
                    schedule.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        # ---- capture -------------------------------------------------------------------------------------------------
        self._zero(params)
        ops.refresh_plain_copies(dev)          # builds the static tables of the one-launch weight re-cast (not capturable)
        ops.bump_param_epoch()                 # derived weight copies are rebuilt INSIDE the graph
        ops.reset_arenas()
        self._defer_prev = trial_parallel.defer_comm if trial_parallel is not None else False
        if multi:
            trial_parallel.defer_comm = True   # hooks pack the buckets; the all-reduces are issued after the replay
        one_graph = optimizer is not None and not multi
        from . import _lib as _L
        n0 = _L.launch_count()
        with torch.cuda.graph(self.graph):
            ops.advance_seed_epoch()
            ops.refresh_plain_copies(dev)      # every bf16 weight copy from its fp32 master: ONE launch per replay
            out = self._fwd_bwd()
            if one_graph:
                optimizer.step()
                if schedule is not None:
                    schedule.step()
            self.outputs = {k: v.detach() for k, v in out.items() if isinstance(v, torch.Tensor)}
        self.captured_launches = _L.launch_count() - n0      # this library's kernels inside one replay
        ops.reset_arenas()
        if trial_parallel is not None:
            trial_parallel.finish()            # (re-arms the buckets; reduces the packed buckets once when sharded)
        if optimizer is not None and multi:
            # the optimizer tail as its own graph, replayed after the all-reduces (gradients are the buckets' flat views)
            self.opt_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.opt_graph, pool=self.graph.pool()):
                optimizer.step()
                if schedule is not None:
                    schedule.step()
            ops.reset_arenas()
        ops.bump_param_epoch()                 # eager code after this must not trust copies derived during capture
        if trial_parallel is not None:
            trial_parallel.defer_comm = self._defer_prev
        # the gradient tensors the graph writes (static addresses): re-attached after every replay, so that eager code in
        # between (which may drop .grad) cannot make them unreachable
        self._params = params
        self._grads = [p.grad for p in params]

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _zero(params) -> None:
        for p in params:
            p.grad = None

    def _fwd_bwd(self) -> Dict[str, torch.Tensor]:
        out = self.loss_fn(self.model, self.batch)
        if not isinstance(out, dict):
            out = {"loss": out}
        out["loss"].backward()
        return out

    # ------------------------------------------------------------------------------------------------------------
    def load(self, batch: Dict[str, torch.Tensor], non_blocking: bool = True) -> None:
        """Refill the static input tensors (device-to-device or pinned-host-to-device copies on the current stream)."""
        for k, v in batch.items():
            self.batch[k].copy_(v, non_blocking=non_blocking)

    def replay(self) -> torch.Tensor:
        """One training step: a single graph launch (plus the gradient all-reduces when sharded).  Returns the static
        loss tensor (device; reading it synchronises)."""
        self.graph.replay()
        for p, g in zip(self._params, self._grads):
            if p.grad is not g:
                p.grad = g
        if self.tp is not None and self.tp.world > 1:
            self.tp.defer_comm = True
            self.tp.finish()
            self.tp.defer_comm = self._defer_prev
            if self.opt_graph is not None:
                self.opt_graph.replay()
        self.replays += 1
        if self.optimizer is not None:
            ops.bump_param_epoch()             # parameters changed behind autograd's version counters
        return self.outputs["loss"]

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        return self.replay()


class GraphedForward:
    """Inference counterpart: ``fn(model, batch)`` (a tensor or a dict of tensors) captured under ``torch.no_grad()``."""

    def __init__(self, model: torch.nn.Module, fn: Callable, example_batch: Dict[str, torch.Tensor], warmup: int = 2):
        dev = next(iter(example_batch.values())).device
        self.model, self.fn = model, fn
        self.batch = {k: v.clone() for k, v in example_batch.items()}
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                fn(model, self.batch)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        ops.bump_param_epoch()
        ops.reset_arenas()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            out = fn(model, self.batch)
            out = out if isinstance(out, dict) else {"out": out}
            self.outputs = {k: v for k, v in out.items() if isinstance(v, torch.Tensor)}
        ops.reset_arenas()
        ops.bump_param_epoch()

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        if batch is not None:
            for k, v in batch.items():
                self.batch[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.outputs
