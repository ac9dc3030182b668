"""Training-step tail of the path: global-norm gradient clipping + AdamW in two kernel
launches over all parameters, with the clip coefficient kept on the device.

Drop-in for the two calls the reference's loops make right after ``loss.backward()``
(``train_art.py:221-229``, ``train_multimodal_fuzzy_fusion.py:464-472``)::

    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)      # ~4 foreach passes
    optimizer.step()                                                  # torch.optim.AdamW: ~6 foreach passes

``FusedClipAdamW`` takes ``torch.optim.AdamW``'s constructor arguments (same defaults, same update rule: decoupled weight
decay, bias-corrected moments) plus ``max_grad_norm``; ``state_dict()`` has AdamW's layout (``step``, ``exp_avg``,
``exp_avg_sq`` per parameter), so checkpoints move both ways.  fp32 CUDA parameters only; no CPU path.
"""
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from . import ops

_CHUNK = 16384                       # elements per CTA


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or weight_decay < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self._plans = {}             # id(group) -> launch plan
        self._sqnorm = None          # device scalar: sum of squared gradients of the last step (all groups)

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
            if plan["copied"] is not None:
                plan["copied"].synchronize()     # the previous step's table upload has left the pinned buffer
            keep = []
            hn, steps = plan["host_np"], plan["steps"]      # numpy views: no tensor objects in the per-parameter loop
            for i, p in enumerate(plan["params"]):
                g = p.grad
                if g is None:
                    hn[i, 1] = 0
                    continue
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                hn[i, 1] = g.data_ptr()
                steps[i] += 1                                # per parameter, like torch: one without gradient lags
            hn[:, 4] = steps
            plan["dev"].copy_(host, non_blocking=True)
            plan["copied"] = torch.cuda.Event()
            plan["copied"].record()
            plans.append((group, plan, keep))
        if not plans:
            return loss
        stream = torch.cuda.current_stream().cuda_stream
        sq = None
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            sq = ops.small_zeros((1,), plans[0][1]["dev"].device)     # the norm is global: over every group
            for _, plan, _ in plans:
                L.call("egb_multi_tensor_sqnorm", plan["dev"].data_ptr(), plan["chunks"].data_ptr(), plan["n_chunks"],
                       sq.data_ptr(), stream)
        self._sqnorm = sq
        for group, plan, _ in plans:
            b1, b2 = group["betas"]
            L.call("egb_multi_tensor_adamw", plan["dev"].data_ptr(), plan["chunks"].data_ptr(), plan["n_chunks"],
                   float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                   float(self.max_grad_norm or 0.0), sq.data_ptr() if sq is not None else None, stream)
        ops.bump_param_epoch()       # the kernels wrote the parameters behind autograd's version counters
        return loss

    def _sync_steps(self, only=None):
        for plan in ([only] if only is not None else self._plans.values()):
            for p, k in zip(plan["params"], plan["steps"]):
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
