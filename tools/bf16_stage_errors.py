"""Where does the bf16 forward error of the EEG model come from?  Runs the CUDA path twice (fp32 parity mode, bf16) on the
same inputs with forward hooks on the sub-modules and prints, per module output, max-abs-err / max-abs and the
Frobenius-relative error of bf16 against fp32."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eyegaze_multimodal_b200.dual_eeg_transformer import DualEEGTransformer  # noqa: E402
from eyegaze_multimodal_b200.precision import precision  # noqa: E402
from eyegaze_multimodal_b200.synth import eeg_pair_batch  # noqa: E402
from oracle import eeg as O  # noqa: E402

DEV = "cuda:0"
cfg, T, B = O.EEGConfig(in_channels=32, max_len=256), 1024, 8
sd = O.init_state_dict(cfg, 2)
m = DualEEGTransformer(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
m.load_state_dict(sd, strict=True)
m = m.to(DEV).eval()
e1, e2 = eeg_pair_batch(B, cfg.in_channels, T, seed=2, coupled=True)
labels = torch.arange(B) % 3
store = {}


def flat(o):
    if torch.is_tensor(o):
        return [o]
    if isinstance(o, (tuple, list)):
        return [t for x in o for t in flat(x)]
    if isinstance(o, dict):
        return [t for x in o.values() for t in flat(x)]
    return []


def hook(name, tag):
    def f(mod, inp, out):
        store.setdefault(name, {})[tag] = [t.detach().float().cpu() for t in flat(out) if t.is_floating_point()]
    return f


for tag in ("fp32", "bf16"):
    hs = [mod.register_forward_hook(hook(n, tag)) for n, mod in m.named_modules() if n and n.count(".") <= 2]
    with precision(tag), torch.no_grad():
        out = m(e1.to(DEV), e2.to(DEV), labels.to(DEV))
    store.setdefault("OUT.logits", {})[tag] = [out["logits"].float().cpu()]
    for h in hs:
        h.remove()
for n, d in store.items():
    if "fp32" not in d or "bf16" not in d:
        continue
    for i, (a, b) in enumerate(zip(d["fp32"], d["bf16"])):
        if a.shape != b.shape or a.numel() == 0:
            continue
        print("%-44s[%d] %-22s max-rel %.3e  fro-rel %.3e" % (n, i, tuple(a.shape), (a - b).abs().max().item() / (a.abs().max().item() + 1e-30),
                                                         (a - b).norm().item() / (a.norm().item() + 1e-30)))
