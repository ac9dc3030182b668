"""Where does the bf16 mode drift from fp32?  Per-submodule relative error (same kernels, two precisions)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer
from eyegaze_multimodal_b200.precision import precision
from eyegaze_multimodal_b200.synth import eeg_pair_batch
from oracle import eeg as O
dev = "cuda:0"
cfg = O.EEGConfig(in_channels=32, max_len=256)
for seed in (2, 5, 7):
    sd = O.init_state_dict(cfg, seed=seed)
    m = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m.load_state_dict(sd); m = m.to(dev).eval()
    e1, e2 = eeg_pair_batch(4, 32, 1024, seed=seed, coupled=True)
    e1, e2 = e1.to(dev), e2.to(dev)
    caps = {}
    def mk(name, store):
        def hook(mod, inp, out):
            t = out[0] if isinstance(out, tuple) else out
            if isinstance(t, torch.Tensor):
                store[name] = t.detach().float().clone()
        return hook
    names = ["ibs_tokenizer", "encoder.layers.0", "encoder.layers.2", "encoder.layers.5", "encoder", "cross_attn.norm"]
    res = {}
    for mode in ("fp32", "bf16"):
        store = {}
        hs = []
        for n in names:
            mod = m.get_submodule(n)
            hs.append(mod.register_forward_hook(mk(n, store)))
        with precision(mode), torch.no_grad():
            out = m(e1, e2)
        for h in hs: h.remove()
        store["logits"] = out["logits"].float()
        store["cls1"] = out["cls1"].float()
        res[mode] = store
    print("seed", seed, " ".join("%s=%.2e" % (n, ((res["bf16"][n] - res["fp32"][n]).abs().max() / res["fp32"][n].abs().max()).item()) for n in res["fp32"]))
    print("        rms:", " ".join("%s=%.2e" % (n, ((res["bf16"][n] - res["fp32"][n]).norm() / res["fp32"][n].norm()).item()) for n in res["fp32"]))
