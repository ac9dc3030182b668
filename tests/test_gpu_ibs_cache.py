"""GPU: the IBS-matrix cache around the real generator -- cached matrices are bit-identical to recomputed ones, the model's
logits do not change, and a warm cache skips the connectivity kernels.  Run on the B200 box:  pytest -m gpu"""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200 import _lib as L
    from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
    from eyegaze_multimodal_b200.ibs_cache import CachedIBSMatrixGenerator
    from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import eeg_pair_batch
from oracle import eeg as O

DEV = "cuda:0"


def test_cached_ibs_matches_recomputation(cuda_device):
    cfg = O.EEGConfig(in_channels=8, d_model=64, num_layers=2, num_heads=4, d_ff=128, max_len=96)
    model = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    model.load_state_dict(O.init_state_dict(cfg, 3), strict=True)
    model = model.to(DEV).eval()
    e1, e2 = eeg_pair_batch(6, 8, 256, seed=4, coupled=True)
    e1, e2 = e1.to(DEV), e2.to(DEV)
    with precision("fp32"), torch.no_grad():
        want = model(e1, e2)["logits"]
        direct = model.ibs_matrix_generator(e1, e2)
        cache = CachedIBSMatrixGenerator(model.ibs_matrix_generator, capacity=32)
        model.ibs_matrix_generator = cache
        keys = [10, 11, 12, 13, 14, 15]
        cache.set_keys(keys)
        cold = model(e1, e2)["logits"]
        n0 = L.launch_count()
        cache.set_keys(keys)
        warm = model(e1, e2)["logits"]
        n_warm = L.launch_count() - n0
        n0 = L.launch_count()
        model(e1, e2)                                   # no keys: transparent, recomputes
        n_plain = L.launch_count() - n0
        cache.set_keys(keys[::-1])                      # another order of the same windows
        rev = model(e1.flip(0), e2.flip(0))["logits"]
        cache.set_keys(keys)
        assert torch.equal(cache(e1, e2), direct)       # bit-identical matrices
    assert torch.equal(cold, want) and torch.equal(warm, want) and torch.equal(rev, want.flip(0))
    assert (cache.hits, cache.misses) == (6 + 6 + 6, 6)
    assert n_warm < n_plain                             # the connectivity kernels did not run on the warm pass
