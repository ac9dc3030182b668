"""Training-step tail of the path: global-norm gradient clipping + AdamW in two kernel
launches over all parameters, with the clip coefficient kept on the device.

Drop-in for the two calls the reference's loops make right after ``loss.backward()``
(``train_art.py:221-229``, ``train_multimodal_fuzzy_fusion.py:464-472``)::

    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)      # ~4 foreach passes
    optimizer.step()                                                  # torch.optim.AdamW: ~6 foreach passes

``FusedClipAdamW`` takes ``torch.optim.AdamW``'s constructor arguments (same defaults, same update rule: decoupled weight
decay, bias-corrected moments) plus ``max_grad_norm``; ``state_dict()`` has AdamW's layout (``step``, ``exp_avg``,
``exp_avg_sq`` per parameter), so checkpoints move both ways.  fp32 CUDA parameters only; no CPU path.

The rest of the training-step tail lives here too:

  * ``torch.amp.GradScaler`` (the reference's multimodal loop trains under fp16 autocast,
    train_multimodal_fuzzy_fusion.py:436-472): the optimizer implements torch's ``_step_supports_amp_scaling`` protocol --
    ``scaler.step(opt)`` hands over ``grad_scale`` / ``found_inf`` as DEVICE tensors and the kernels un-scale, clip and skip
    on the device, so the scaler's per-step ``.item()`` on found_inf disappears.  Independently, ``skip_nonfinite=True``
    skips an update whose global gradient norm (our own norm pass) is inf / nan.
  * ``capturable=True``: learning rate and step count live in device memory (``DeviceLRSchedule`` advances them), the
    gradient pointer table is static, so ``step()`` enqueues identical launches every time and can sit inside a captured
    CUDA graph (graphs.GraphedTrainStep).
  * ``DeviceLRSchedule``: CosineAnnealingLR (train_art.py:401-409) and linear-warm-up + cosine LambdaLR
    (train_multimodal_fuzzy_fusion.py:197-214) evaluated by a one-thread kernel; per-group base learning rates
    (:727-736) are kept.
  * ``MetricAccumulator``: running sums of loss scalars and the argmax-accuracy count on the device; ONE host read per
    epoch instead of six ``.item()`` calls per step (train_art.py:224-229).
"""
import ctypes as C
import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops
from . import torch_ops as TO

_CHUNK = 16384                       # elements per CTA


class FusedClipAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True    # torch.amp.GradScaler.step() then passes grad_scale / found_inf as device tensors

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, skip_nonfinite: bool = False, capturable: bool = False):
        if lr < 0 or eps < 0 or weight_decay < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.skip_nonfinite = skip_nonfinite
        self.capturable = capturable
        self._plans = {}             # id(group) -> launch plan
        self._sqnorm = None          # device scalar: sum of squared gradients of the last step (all groups)
        self._dev_state = None       # capturable: [opt_step, lr_0 .. lr_{G-1}] fp32 on the device
        self._sq_static = None       # capturable: the norm accumulator has a fixed address
        # (GradScaler.step() sets / deletes the attributes `grad_scale` and `found_inf` around step(); they must not exist
        # in between: torch multiplies its scale with getattr(optimizer, "grad_scale", 1))

    # -- device-resident step state --------------------------------------------------------------------
    def device_state(self) -> torch.Tensor:
        """fp32 device vector [skip flag, steps skipped, step count, -, lr of group 0, lr of group 1, ...].  The first four
        words are written by egb_adamw_prepare on every step that needs a device-side verdict (GradScaler / non-finite
        skip / capturable mode); the learning rates are read by the update kernel in capturable mode and rewritten by
        ``DeviceLRSchedule``.  Created on first use from the groups' current ``lr``."""
        if self._dev_state is None:
            dev = next(p for g in self.param_groups for p in g["params"]).device
            st = torch.zeros(4 + len(self.param_groups), dtype=torch.float32)
            for i, g in enumerate(self.param_groups):
                st[4 + i] = float(g["lr"])
            self._dev_state = st.to(dev)
        return self._dev_state

    def device_step(self) -> torch.Tensor:
        """Capturable mode: the 1-based count of updates actually applied (device scalar view)."""
        return self.device_state()[2]

    def device_lrs(self) -> torch.Tensor:
        return self.device_state()[4:]

    # ------------------------------------------------------------------------------------------------
    def _plan(self, group):
        plan = self._plans.get(id(group))
        params = [p for p in group["params"] if p.requires_grad]
        # a plan caches raw addresses: it is valid for the same parameter objects AT the same storage only
        # (model.to(...) / p.data = ... swap the storage under an unchanged id)
        key = [(id(p), p.data_ptr()) for p in params]
        if plan is not None and plan["ids"] == key:
            return plan
        if plan is not None:
            self._sync_steps(plan)   # the live step counts move to self.state before the plan is rebuilt
        if not params:
            self._plans.pop(id(group), None)
            return None
        dev = params[0].device
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise TypeError("FusedClipAdamW expects contiguous fp32 CUDA parameters on one device")
        # moments live in two flat buffers; the per-parameter state tensors are views (AdamW's state_dict layout)
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        m_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        v_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            st = self.state[p]
            ea = m_flat[o:o + p.numel()].view_as(p)
            es = v_flat[o:o + p.numel()].view_as(p)
            if "exp_avg" in st:      # restored from a checkpoint (or a previous plan)
                ea.copy_(st["exp_avg"])
                es.copy_(st["exp_avg_sq"])
            st["exp_avg"], st["exp_avg_sq"] = ea, es
            st.setdefault("step", torch.tensor(0.0))
        steps = np.array([int(self.state[p]["step"]) for p in params], dtype=np.int64)
        # tensor table (pinned host copy; the gradient column is refreshed every step) and the static chunk table
        host = torch.zeros(len(params), 5, dtype=torch.int64).pin_memory()   # p, g, exp_avg, exp_avg_sq, step
        for i, p in enumerate(params):
            host[i, 0] = p.data_ptr()
            host[i, 2] = self.state[p]["exp_avg"].data_ptr()
            host[i, 3] = self.state[p]["exp_avg_sq"].data_ptr()
        chunks = []
        for i, p in enumerate(params):
            for o in range(0, p.numel(), _CHUNK):
                chunks.append((i, min(_CHUNK, p.numel() - o), o))
        ch = np.zeros(len(chunks), dtype=[("tensor", "<i4"), ("n", "<i4"), ("offset", "<i8")])
        ch["tensor"], ch["n"], ch["offset"] = zip(*chunks)
        plan = dict(ids=key, params=params, host=host, dev=torch.empty_like(host, device=dev),
                    chunks=torch.from_numpy(ch.view(np.uint8).copy()).to(dev), n_chunks=len(chunks), m=m_flat, v=v_flat,
                    copied=None, steps=steps, host_np=host.numpy())
        self._plans[id(group)] = plan
        return plan

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plans = []
        for group in self.param_groups:
            plan = self._plan(group)
            if plan is None:
                continue
            host = plan["host"]
            if plan["copied"] is not None and not self.capturable:
                plan["copied"].synchronize()     # the previous step's table upload has left the pinned buffer
            keep = []
            hn, steps = plan["host_np"], plan["steps"]      # numpy views: no tensor objects in the per-parameter loop
            changed = False
            for i, p in enumerate(plan["params"]):
                g = p.grad
                if g is None:
                    changed |= hn[i, 1] != 0
                    hn[i, 1] = 0
                    continue
                if g.dtype != torch.float32 or not g.is_contiguous():
                    if self.capturable:
                        raise TypeError("capturable FusedClipAdamW needs contiguous fp32 gradients")
                    g = g.float().contiguous()
                    keep.append(g)
                changed |= hn[i, 1] != g.data_ptr()
                hn[i, 1] = g.data_ptr()
                steps[i] += 1                                # per parameter, like torch: one without gradient lags
            hn[:, 4] = steps
            if self.capturable:
                # the table is static (gradients live at fixed addresses, the step count is on the device): upload it
                # only when it changed, never while a graph is being captured
                # (inside a capture this becomes a memcpy node that re-uploads the same 40-byte records on each replay)
                if changed or plan["copied"] is None:
                    plan["dev"].copy_(host, non_blocking=True)
                    plan["copied"] = True
            else:
                plan["dev"].copy_(host, non_blocking=True)
                plan["copied"] = torch.cuda.Event()
                plan["copied"].record()
            plans.append((group, plan, keep))
        if not plans:
            return loss
        dev = plans[0][1]["dev"].device
        sq = None
        if (self.max_grad_norm is not None and self.max_grad_norm > 0) or self.skip_nonfinite:
            if self.capturable:
                if self._sq_static is None:
                    self._sq_static = torch.zeros(1, dtype=torch.float32, device=dev)
                sq = self._sq_static
                TO.call("zero", sq, 4)
            else:
                sq = ops.small_zeros((1,), dev)                       # the norm is global: over every group
            for _, plan, _ in plans:
                TO.call("multi_tensor_sqnorm", plan["dev"], plan["chunks"], plan["n_chunks"],
                       sq)
        self._sqnorm = sq
        st = TO.AdamwState()
        gs, fi = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        keep_amp = []
        if gs is not None:
            gs = gs.to(device=dev, dtype=torch.float32).reshape(1)
            st.grad_scale = gs
            keep_amp.append(gs)
        if fi is not None:
            fi = fi.to(device=dev, dtype=torch.float32).reshape(1)
            keep_amp.append(fi)
        needs_verdict = self.capturable or self.skip_nonfinite or fi is not None or self._dev_state is not None
        dstate = self.device_state() if needs_verdict else None
        if dstate is not None:
            # one thread decides whether this step is skipped and keeps the skipped / applied counts on the device
            st.ctrl = dstate
            st.use_device_step = 1 if self.capturable else 0
            TO.call("adamw_prepare", dstate, sq if sq is not None else None,
                   gs if gs is not None else None, fi if fi is not None else None,
                   1 if self.skip_nonfinite else 0, 1)
        for gi, (group, plan, _) in enumerate(plans):
            b1, b2 = group["betas"]
            if self.capturable:
                st.lr = dstate[4 + self.param_groups.index(group):]
            TO.call("multi_tensor_adamw_ex", plan["dev"], plan["chunks"], plan["n_chunks"],
                   float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                   float(self.max_grad_norm or 0.0), sq if sq is not None else None, st)
        ops.bump_param_epoch()       # the kernels wrote the parameters behind autograd's version counters
        if not torch.cuda.is_current_stream_capturing():
            ops.refresh_plain_copies()   # eager loops: all bf16 weight copies re-derived now, in one launch
        return loss

    def _sync_steps(self, only=None):
        dev_step = skipped = None
        if self._dev_state is not None:
            ds = self._dev_state[:3].tolist()
            skipped = int(ds[1])                            # steps the device skipped (GradScaler / non-finite norm)
            if self.capturable:
                dev_step = ds[2]                            # graph replays advance the count on the device only
        for plan in ([only] if only is not None else self._plans.values()):
            for i, (p, k) in enumerate(zip(plan["params"], plan["steps"])):
                if dev_step is not None and plan["host_np"][i, 1] != 0:
                    k = dev_step
                    plan["steps"][i] = int(dev_step)
                elif skipped:
                    k = max(int(k) - skipped, 0)
                self.state[p]["step"] = torch.tensor(float(k))

    def state_dict(self):
        self._sync_steps()           # the per-parameter step counts live in a host array between checkpoints
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch keeps tensors that already have the parameter's dtype and device BY REFERENCE: detach from the donor
        # optimizer's flat buffers now, the packing itself happens lazily on the next step
        for st in self.state.values():
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st:
                    st[k] = st[k].clone()
        self._plans.clear()          # moments and step counts are re-packed on the next step

    def grad_norm(self) -> Optional[torch.Tensor]:
        """Total gradient norm of the last step as a DEVICE scalar (what clip_grad_norm_ returns); no host sync."""
        return None if self._sqnorm is None else self._sqnorm.sqrt().squeeze(0)


class DeviceLRSchedule:
    """Learning-rate schedule evaluated on the device for a ``FusedClipAdamW(capturable=True)``.

    ``kind='cosine'``        : torch.optim.lr_scheduler.CosineAnnealingLR(T_max, eta_min) in closed form, stepped once per
                               EPOCH by the reference (train_art.py:401-409, 445);
    ``kind='warmup_cosine'`` : get_linear_warmup_cosine_scheduler (train_multimodal_fuzzy_fusion.py:197-214), a LambdaLR
                               stepped once per optimizer STEP (:503-504);
    ``kind='constant'``.
    Like torch's schedulers, construction sets the learning rates for step 0; ``step()`` advances the counter by one and
    rewrites every group's rate -- a one-thread kernel, no host value involved, so it can be captured in a CUDA graph.
    """
    KINDS = {"constant": 0, "cosine": 1, "warmup_cosine": 2}

    def __init__(self, optimizer: FusedClipAdamW, kind: str = "constant", T_max: float = 1.0, eta_min: float = 0.0,
                 warmup_steps: float = 0.0, total_steps: float = 1.0):
        if kind not in self.KINDS:
            raise ValueError("kind must be one of %s" % sorted(self.KINDS))
        if not optimizer.capturable:
            raise ValueError("DeviceLRSchedule drives a FusedClipAdamW(capturable=True)")
        self.opt, self.kind = optimizer, kind
        self.p0, self.p1 = (float(T_max), float(eta_min)) if kind == "cosine" else (float(warmup_steps), float(total_steps))
        st = optimizer.device_state()
        dev = st.device
        self.base_lrs = [float(g.setdefault("initial_lr", g["lr"])) for g in optimizer.param_groups]
        self._base = torch.tensor(self.base_lrs, dtype=torch.float32, device=dev)
        self._sched = torch.zeros(1, dtype=torch.float32, device=dev)        # the scheduler's own counter
        self._launch(advance=0)

    def _launch(self, advance: int) -> None:
        st = self.opt.device_state()
        TO.call("lr_schedule_step", self._sched, None, self._base, st[4:],
               len(self.base_lrs), self.KINDS[self.kind], self.p0, self.p1, advance, 0)

    def step(self) -> None:
        self._launch(advance=1)

    def get_last_lr(self) -> Sequence[float]:
        """Host read (synchronises): for logging only."""
        return self.opt.device_lrs().tolist()

    def state_dict(self) -> Dict:
        return {"kind": self.kind, "p0": self.p0, "p1": self.p1, "last_epoch": float(self._sched.item()),
                "opt_step": float(self.opt.device_step().item()), "base_lrs": list(self.base_lrs)}

    def load_state_dict(self, sd: Dict) -> None:
        self.kind, self.p0, self.p1 = sd["kind"], sd["p0"], sd["p1"]
        self._sched.fill_(sd["last_epoch"])
        self.opt.device_state()[2] = sd["opt_step"]
        self._launch(advance=0)

    @staticmethod
    def reference_factor(kind: str, t: float, p0: float, p1: float) -> float:
        """The closed forms above on the host (documentation / tests)."""
        if kind == "cosine":
            return 0.5 * (1.0 + math.cos(math.pi * t / p0))
        if kind == "warmup_cosine":
            if t < p0:
                return t / max(1.0, p0)
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * (t - p0) / max(1.0, p1 - p0))))
        return 1.0


class MetricAccumulator:
    """Running sums of per-step loss scalars and the argmax accuracy, kept on the device.

    The reference's loops read every loss with ``.item()`` every step (train_art.py:224-229: six synchronisations per
    step; train_multimodal_fuzzy_fusion.py:507-517: predictions, labels, alphas and five losses) -- each one drains the
    GPU.  ``add(**scalars)`` / ``add_predictions(logits, labels)`` enqueue one tiny kernel each; ``result()`` does ONE host
    read (per epoch, or whenever the loop wants to print)."""

    def __init__(self, names: Sequence[str], device):
        if not 0 < len(names) <= 8:
            raise ValueError("1..8 named scalars")
        self.names = list(names)
        self._acc = torch.zeros(len(names) + 1 + 2, dtype=torch.float32, device=device)   # sums | batches | hits, trials
        self._zero = torch.zeros((), dtype=torch.float32, device=device)

    def add(self, **scalars: torch.Tensor) -> None:
        keep = []
        for i, n in enumerate(self.names):
            t = scalars.get(n, self._zero)
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_cuda:
                t = t.float().to(self._acc.device)
            t = t.reshape(())
            keep.append(t)
        TO.call("accum_scalars", keep, len(self.names), self._acc)

    def add_predictions(self, logits: torch.Tensor, labels: torch.Tensor, preds_out: Optional[torch.Tensor] = None) -> None:
        logits = logits.detach().float().contiguous()
        labels = labels.contiguous().long()
        TO.call("argmax_count", logits, labels, self._acc[len(self.names) + 1:],
               preds_out if preds_out is not None else None, logits.shape[0], logits.shape[1])

    def result(self) -> Dict[str, float]:
        a = self._acc.tolist()                          # the single host read
        n = len(self.names)
        batches = max(a[n], 1.0)
        out = {k: a[i] / batches for i, k in enumerate(self.names)}
        out["batches"] = a[n]
        if a[n + 2] > 0:
            out["accuracy"] = a[n + 1] / a[n + 2]
        return out

    def reset(self) -> None:
        self._acc.zero_()
