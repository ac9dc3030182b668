"""Device-resident cache of inter-brain-synchrony matrices keyed by dataset window.

``IBSConnectivityMatrixGenerator`` (dual_eeg_transformer.py:473-819) has no parameters and is deterministic: the
(6, F, C, C) matrices of a window never change over training, yet the reference -- and the path as benchmarked -- recompute
them every epoch (4 ms of the 53 ms cfg2 step).  ``CachedIBSMatrixGenerator`` wraps the generator module; the training
loop names the windows of the next batch with ``set_keys`` (e.g. the dataset indices it already has), hits are gathered
from a preallocated device buffer, misses are computed by the wrapped generator and stored.  172 KB per window at
C = 32: ten thousand windows take 1.7 GB of the 180 GB.

    model.ibs_matrix_generator = CachedIBSMatrixGenerator(model.ibs_matrix_generator, capacity=len(dataset))
    for idx, batch in loader:                      # idx: the windows' dataset indices
        model.ibs_matrix_generator.set_keys(idx)
        out = model(batch['eeg1'], batch['eeg2'], batch['labels'])

Without ``set_keys`` the wrapper is transparent.  Forward hooks that the reference's analysis code registers on
``model.ibs_matrix_generator`` (5_Metrics/eeg_metrics.py:203,341) keep working: they now sit on the wrapper and still see
and may replace the matrices.  The cache holds no autograd state (the matrices do not depend on parameters; gradients
with respect to the EEG input are not defined by the reference either: det:593 runs under no_grad)."""
from typing import Iterable, Optional

import torch
import torch.nn as nn


class CachedIBSMatrixGenerator(nn.Module):
    def __init__(self, generator: nn.Module, capacity: int):
        super().__init__()
        if capacity < 1:
            raise ValueError("capacity must be positive")
        self.generator = generator
        self.capacity = int(capacity)
        self._slot_of = {}             # key -> slot
        self._key_of = [None] * self.capacity
        self._next = 0                 # ring pointer: the oldest entry is evicted first
        self._store = None             # (capacity, *item) on the generator's output device
        self._keys = None
        self.hits = 0
        self.misses = 0

    # the wrapped module's public attributes stay reachable (num_features, band_names, ...)
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("generator"), name)

    def set_keys(self, keys: Optional[Iterable[int]]) -> None:
        """Keys (one hashable per trial, in batch order) of the NEXT forward; ``None`` bypasses the cache once."""
        if keys is None:
            self._keys = None
        elif torch.is_tensor(keys):
            self._keys = [int(k) for k in keys.reshape(-1).tolist()]
        else:
            self._keys = list(keys)

    def clear(self) -> None:
        self._slot_of.clear()
        self._key_of = [None] * self.capacity
        self._next = 0
        self.hits = self.misses = 0

    @torch.no_grad()
    def forward(self, eeg1: torch.Tensor, eeg2: torch.Tensor) -> torch.Tensor:
        keys, self._keys = self._keys, None
        if keys is None:
            return self.generator(eeg1, eeg2)
        B = eeg1.shape[0]
        if len(keys) != B:
            raise ValueError("set_keys got %d keys for a batch of %d trials" % (len(keys), B))
        # a key that appears twice in the batch is computed once; more distinct misses than the cache can hold at once
        # cannot be served from it
        miss_rows, miss_keys, seen = [], [], set()
        for i, k in enumerate(keys):
            if k not in self._slot_of and k not in seen:
                seen.add(k)
                miss_rows.append(i)
                miss_keys.append(k)
        if len(miss_keys) > self.capacity or len(set(keys)) > self.capacity:
            return self.generator(eeg1, eeg2)
        self.misses += len(miss_keys)
        self.hits += B - len(miss_keys)
        if miss_rows:
            idx = torch.as_tensor(miss_rows, device=eeg1.device)
            fresh = self.generator(eeg1.index_select(0, idx), eeg2.index_select(0, idx))
            if self._store is None:
                self._store = torch.empty((self.capacity,) + tuple(fresh.shape[1:]), dtype=fresh.dtype, device=fresh.device)
            protected = set(keys)
            slots = []
            for k in miss_keys:
                # ring eviction that never throws out an entry this very batch needs
                while self._key_of[self._next] is not None and self._key_of[self._next] in protected \
                        and self._key_of[self._next] in self._slot_of:
                    self._next = (self._next + 1) % self.capacity
                s = self._next
                old = self._key_of[s]
                if old is not None:
                    self._slot_of.pop(old, None)
                self._key_of[s] = k
                self._slot_of[k] = s
                slots.append(s)
                self._next = (self._next + 1) % self.capacity
            self._store.index_copy_(0, torch.as_tensor(slots, device=fresh.device), fresh)
        gather = torch.as_tensor([self._slot_of[k] for k in keys], device=self._store.device)
        return self._store.index_select(0, gather)
