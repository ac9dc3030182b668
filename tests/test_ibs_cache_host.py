"""CPU: bookkeeping of the IBS-matrix cache (hits, misses, duplicates, ring eviction, bypass) with a stand-in generator --
the cache is plain tensor indexing around whatever module it wraps; the real generator is exercised in test_gpu_model."""
import torch
import torch.nn as nn

from eyegaze_multimodal_b200.ibs_cache import CachedIBSMatrixGenerator


class FakeGenerator(nn.Module):
    num_features = 7

    def __init__(self):
        super().__init__()
        self.rows_computed = 0

    def forward(self, a, b):
        self.rows_computed += a.shape[0]
        return torch.stack([a.sum((1, 2)), b.sum((1, 2)), (a * b).sum((1, 2))], 1).view(-1, 3, 1)


def _batch(keys):
    g = torch.Generator().manual_seed(0)
    table = torch.randn(100, 2, 4, 8, generator=g)
    return table[keys, 0], table[keys, 1]


def test_hits_misses_duplicates_and_bypass():
    gen = FakeGenerator()
    c = CachedIBSMatrixGenerator(gen, capacity=16)
    assert c.num_features == 7                                   # attributes of the wrapped module stay reachable
    k1 = [3, 5, 3, 9]                                            # a duplicate inside the batch is computed once
    c.set_keys(k1)
    out = c(*_batch(k1))
    assert torch.equal(out, FakeGenerator()(*_batch(k1))) and gen.rows_computed == 3
    k2 = torch.tensor([9, 5, 11, 3])
    c.set_keys(k2)
    out = c(*_batch(k2.tolist()))
    assert torch.equal(out, FakeGenerator()(*_batch(k2.tolist()))) and gen.rows_computed == 4   # only key 11 was new
    assert (c.hits, c.misses) == (1 + 3, 3 + 1)
    out = c(*_batch([1, 2]))                                     # no keys announced: transparent
    assert gen.rows_computed == 6 and out.shape[0] == 2
    hooked = []
    h = c.register_forward_hook(lambda m, i, o: hooked.append(o.shape) or o * 0)   # analysis hooks sit on the wrapper
    c.set_keys([3])
    assert torch.count_nonzero(c(*_batch([3]))) == 0 and hooked
    h.remove()


def test_ring_eviction_keeps_the_current_batch():
    gen = FakeGenerator()
    c = CachedIBSMatrixGenerator(gen, capacity=4)
    for keys in ([0, 1, 2, 3], [4, 5, 0, 1], [2, 3, 6, 7], [6, 7, 6, 7]):
        c.set_keys(keys)
        assert torch.equal(c(*_batch(keys)), FakeGenerator()(*_batch(keys))), keys
    assert len(c._slot_of) <= 4
    c.set_keys(list(range(10)))                                  # more distinct windows than the cache holds: computed directly
    assert torch.equal(c(*_batch(list(range(10)))), FakeGenerator()(*_batch(list(range(10)))))
    c.clear()
    assert c.hits == 0 and not c._slot_of
