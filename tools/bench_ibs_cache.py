"""cfg2 geometry: IBS matrices of 256 trials (32 ch x 1024) computed vs served from the cache."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200.dual_eeg_transformer import IBSConnectivityMatrixGenerator
from eyegaze_multimodal_b200.ibs_cache import CachedIBSMatrixGenerator
dev = "cuda:0"
B = 256
e1, e2 = torch.randn(B, 32, 1024, device=dev), torch.randn(B, 32, 1024, device=dev)
gen = IBSConnectivityMatrixGenerator(32).to(dev)
cache = CachedIBSMatrixGenerator(gen, capacity=4096)
keys = list(range(B))


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1_.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1_) / iters


def cached():
    cache.set_keys(keys)
    return cache(e1, e2)


t_compute = timeit(lambda: gen(e1, e2))
t_cached = timeit(cached)
item = gen(e1[:1], e2[:1]).numel() * 4
print(json.dumps({"trials": B, "bytes_per_trial": item, "compute_ms": t_compute, "cached_ms": t_cached,
                  "cached_GBps": 2 * B * item / t_cached / 1e6}))
