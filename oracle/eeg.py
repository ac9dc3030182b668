"""Functional fp32 CPU restatement of the dual-EEG encoder (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function takes a ``state_dict``-style mapping ``sd`` (reference key names, Appendix C of SURVEY.md)
and plain tensors; nothing here is an ``nn.Module``.  Reference locations are cited as
``det:<line>`` = 3_Models/backbones/dual_eeg_transformer.py and ``art:<line>`` = 3_Models/backbones/art.py.

The IBS connectivity generator is restated in *vectorised* form (broadcast over channel pairs) so
that it finishes in seconds; the reference loops over (i, j) in Python (det:604-756).
"""
import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# det:500-509 -- band edges in Hz, order fixed
IBS_BANDS = [(0.5, 45.0), (0.5, 4.0), (4.0, 8.0), (8.0, 13.0), (13.0, 30.0), (30.0, 45.0)]
SCALAR_BANDS = [(4.0, 8.0), (8.0, 13.0), (13.0, 30.0), (30.0, 45.0)]  # det:201-206
FEATURE_SUBSETS = {"all": [0, 1, 2, 3, 4, 5, 6], "phase": [0, 1, 2, 5], "amplitude": [3, 4, 6]}  # det:515-525


@dataclass
class EEGConfig:
    """Constructor arguments of DualEEGTransformer (det:995-1021), same names and defaults."""
    in_channels: int = 62
    num_classes: int = 3
    d_model: int = 256
    num_layers: int = 6
    num_heads: int = 8
    d_ff: int = 1024
    dropout: float = 0.1
    max_len: int = 2048
    conv_kernel_size: int = 25
    conv_stride: int = 4
    conv_layers: int = 2
    sampling_rate: int = 256
    use_spectrogram: bool = True
    spec_n_fft: int = 128
    spec_hop_length: int = 64
    spec_freq_bins: int = 64
    use_robust_ibs: bool = True
    use_ibs: bool = True
    use_cross_attention: bool = True
    ibs_instance_norm: bool = True
    ibs_feature_type: str = "all"

    @property
    def num_ibs_features(self) -> int:
        return {"all": 7, "phase": 4, "amplitude": 3}.get(self.ibs_feature_type, 7)  # det:1035-1036

    @property
    def num_ibs_tokens(self) -> int:
        if not self.use_ibs:
            return 0
        return 6 * self.num_ibs_features if self.use_robust_ibs else 1  # det:1037-1043


# ------------------------------------------------------------------------------------------------
# signal prologue shared by both IBS generators
# ------------------------------------------------------------------------------------------------
def bandpass_fft(x: Tensor, lo: float, hi: float, fs: float) -> Tensor:
    """det:527-560: rfft -> inclusive frequency mask -> irfft."""
    T = x.shape[-1]
    freqs = torch.fft.rfftfreq(T, d=1.0 / fs)                      # evaluated on the host, like the reference (det:548)
    keep = ((freqs >= lo) & (freqs <= hi)).to(dtype=x.dtype, device=x.device)
    return torch.fft.irfft(torch.fft.rfft(x, dim=-1) * keep, n=T, dim=-1)


def hilbert_phase(x: Tensor) -> Tensor:
    """det:562-591: analytic signal through the one-sided spectrum weights, then angle."""
    T = x.shape[-1]
    h = torch.zeros(T, dtype=x.dtype, device=x.device)
    if T % 2 == 0:
        h[0] = 1
        h[T // 2] = 1
        h[1:T // 2] = 2
    else:
        h[0] = 1
        h[1:(T + 1) // 2] = 2
    return torch.angle(torch.fft.ifft(torch.fft.fft(x, dim=-1) * h, dim=-1))


def _pearson_rows(a: Tensor, b: Tensor) -> Tensor:
    """det:701-712 / 745-756: z-score with UNBIASED std (+1e-8), mean of products. a:(B,C,T) b:(B,C,T) -> (B,C,C)."""
    an = (a - a.mean(-1, keepdim=True)) / (a.std(-1, keepdim=True) + 1e-8)
    bn = (b - b.mean(-1, keepdim=True)) / (b.std(-1, keepdim=True) + 1e-8)
    return (an.unsqueeze(2) * bn.unsqueeze(1)).mean(-1)


IBS_CHUNK = 4          # trials per vectorised block (bench.py raises it for the eager-on-GPU baseline leg)


def ibs_connectivity(eeg1: Tensor, eeg2: Tensor, fs: float = 256.0, feature_type: str = "all",
                     chunk: Optional[int] = None) -> Tensor:
    """IBSConnectivityMatrixGenerator.forward (det:760-819) -> (B, 6, F, C, C), fp32."""
    chunk = chunk or IBS_CHUNK
    B, C, T = eeg1.shape
    out = torch.zeros(B, 6, 7, C, C, dtype=torch.float32, device=eeg1.device)
    for b0 in range(0, B, chunk):
        x1, x2 = eeg1[b0:b0 + chunk], eeg2[b0:b0 + chunk]
        for bi, (lo, hi) in enumerate(IBS_BANDS):
            a, b = bandpass_fft(x1, lo, hi, fs), bandpass_fft(x2, lo, hi, fs)
            p1, p2 = a ** 2, b ** 2
            ph1, ph2 = hilbert_phase(a), hilbert_phase(b)
            d = ph1.unsqueeze(2) - ph2.unsqueeze(1)                       # (b,C,C,T) raw, un-wrapped
            sgn = torch.sign(d)
            plv = torch.abs(torch.mean(torch.exp(1j * d), dim=-1))       # det:606-609
            pli = torch.abs(sgn.mean(-1))                                  # det:626-628
            w = (p1.unsqueeze(2) + p2.unsqueeze(1)) / 2                    # det:653-656
            w = w / (w.sum(-1, keepdim=True) + 1e-8)
            wpli = torch.abs((sgn * w).sum(-1))
            f1, f2 = torch.fft.rfft(a, dim=-1), torch.fft.rfft(b, dim=-1)  # det:672-686
            pxy = f1.unsqueeze(2) * f2.unsqueeze(1).conj()
            pxx = (f1 * f1.conj()).real.unsqueeze(2)
            pyy = (f2 * f2.conj()).real.unsqueeze(1)
            coh = ((pxy.abs() ** 2) / (pxx * pyy + 1e-8)).mean(-1)
            pcorr = _pearson_rows(p1, p2)                                  # det:701-712
            pdiff = torch.abs(d).mean(-1)                                  # det:729-730
            tcorr = _pearson_rows(a, b)                                    # det:745-756
            for fi, m in enumerate((plv, pli, wpli, coh, pcorr, pdiff, tcorr)):
                out[b0:b0 + chunk, bi, fi] = m.to(torch.float32)
    idx = FEATURE_SUBSETS.get(feature_type, FEATURE_SUBSETS["all"])
    return out[:, :, idx]


def ibs_scalar_features(eeg1: Tensor, eeg2: Tensor, fs: float = 256.0) -> Tensor:
    """IBSTokenGenerator feature stage (det:432-461): 4 bands x 7 global scalars -> (B, 28)."""
    feats = []
    for lo, hi in SCALAR_BANDS:
        a, b = bandpass_fft(eeg1, lo, hi, fs), bandpass_fft(eeg2, lo, hi, fs)
        p1, p2 = a ** 2, b ** 2
        ph1, ph2 = hilbert_phase(a), hilbert_phase(b)
        d = ph1 - ph2
        plv = torch.abs(torch.mean(torch.exp(1j * d), dim=(1, 2)))                      # det:267-271
        pli = torch.abs(torch.sign(d).mean(dim=(1, 2)))                                 # det:334-336
        w = (p1 + p2) / 2
        w = w / (w.sum(dim=(1, 2), keepdim=True) + 1e-8)
        wpli = torch.abs((torch.sign(d) * w).sum(dim=(1, 2)))                           # det:355-365
        f1, f2 = torch.fft.rfft(a, dim=2), torch.fft.rfft(b, dim=2)                     # det:378-394
        pxy = (f1 * f2.conj()).mean(1)
        pxx = (f1 * f1.conj()).mean(1).real
        pyy = (f2 * f2.conj()).mean(1).real
        coh = ((pxy.abs() ** 2) / (pxx * pyy + 1e-8)).mean(1)
        q1, q2 = p1.flatten(1), p2.flatten(1)                                           # det:281-290
        q1 = (q1 - q1.mean(1, keepdim=True)) / (q1.std(1, keepdim=True) + 1e-8)
        q2 = (q2 - q2.mean(1, keepdim=True)) / (q2.std(1, keepdim=True) + 1e-8)
        pcorr = (q1 * q2).mean(1)
        pdiff = torch.abs(torch.mean(d, dim=(1, 2)))                                    # det:455
        m1, m2 = a.mean(1), b.mean(1)                                                   # det:406-416
        m1 = (m1 - m1.mean(1, keepdim=True)) / (m1.std(1, keepdim=True) + 1e-8)
        m2 = (m2 - m2.mean(1, keepdim=True)) / (m2.std(1, keepdim=True) + 1e-8)
        tcorr = (m1 * m2).mean(1)
        feats += [plv, pli, wpli, coh, pcorr, pdiff, tcorr]
    return torch.stack(feats, dim=1).to(torch.float32)


# ------------------------------------------------------------------------------------------------
# sub-modules
# ------------------------------------------------------------------------------------------------
def temporal_conv(x: Tensor, sd: Dict[str, Tensor], cfg: EEGConfig, pre: str = "temporal_conv.") -> Tensor:
    """TemporalConvFrontend (det:163-175), eval mode: (B,C,T) -> (B,T~,d)."""
    for i in range(cfg.conv_layers):
        x = F.relu(F.conv1d(x, sd[f"{pre}convs.{i}.weight"], sd[f"{pre}convs.{i}.bias"], stride=cfg.conv_stride,
                            padding=cfg.conv_kernel_size // 2))
    return x.permute(0, 2, 1)


def spectrogram_logmag(x: Tensor, window: Tensor, cfg: EEGConfig) -> Tensor:
    """det:98-121: STFT (centre, reflect) -> |.| -> first bins -> log(+1e-8): (B,C,T) -> (B*C,1,F,frames)."""
    B, C, T = x.shape
    st = torch.stft(x.reshape(B * C, T), n_fft=cfg.spec_n_fft, hop_length=cfg.spec_hop_length, window=window,
                    return_complex=True, center=True)
    mag = torch.abs(st)[:, :cfg.spec_freq_bins, :]
    return torch.log(mag + 1e-8).unsqueeze(1)


def spectrogram_tokens(x: Tensor, sd: Dict[str, Tensor], cfg: EEGConfig, pre: str = "spectrogram_generator.",
                       taps: Optional[dict] = None) -> Tensor:
    """SpectrogramTokenGenerator.forward (det:88-135), eval mode.  ``taps['spec_conv3']`` collects the spec_conv[3] output
    (what a Grad-CAM hook on that module sees, 5_Metrics/eeg_metrics.py:841), with its gradient retained."""
    B, C, _ = x.shape
    img = spectrogram_logmag(x, sd[pre + "window"], cfg)
    h = F.relu(F.conv2d(img, sd[pre + "spec_conv.0.weight"], sd[pre + "spec_conv.0.bias"], padding=1))
    h = F.max_pool2d(h, 2)
    h = F.conv2d(h, sd[pre + "spec_conv.3.weight"], sd[pre + "spec_conv.3.bias"], padding=1)
    if taps is not None:
        if h.requires_grad:
            h.retain_grad()
        taps.setdefault("spec_conv3", []).append(h)
    h = F.relu(h)
    h = F.adaptive_avg_pool2d(h, (4, 4)).flatten(1)
    h = F.relu(F.linear(h, sd[pre + "proj.0.weight"], sd[pre + "proj.0.bias"]))
    h = F.linear(h, sd[pre + "proj.3.weight"], sd[pre + "proj.3.bias"])
    return h.reshape(B, C, cfg.d_model)


def ibs_tokenize(mats: Tensor, sd: Dict[str, Tensor], cfg: EEGConfig, pre: str = "ibs_tokenizer.") -> Tensor:
    """RobustIBSTokenizer.forward (det:879-911), eval mode: (B,6,F,C,C) -> (B,6F,d)."""
    B, nb, nf, C1, C2 = mats.shape
    assert C1 == C2 == cfg.in_channels and nb == 6 and nf == cfg.num_ibs_features
    x = mats.reshape(B, nb * nf, C1 * C2)
    if cfg.ibs_instance_norm:
        x = F.instance_norm(x.permute(0, 2, 1), weight=sd[pre + "instance_norm.weight"],
                            bias=sd[pre + "instance_norm.bias"], eps=1e-5).permute(0, 2, 1)
    x = F.gelu(F.linear(x, sd[pre + "bottleneck.0.weight"], sd[pre + "bottleneck.0.bias"]))
    x = F.linear(x, sd[pre + "bottleneck.3.weight"], sd[pre + "bottleneck.3.bias"])
    return x + sd[pre + "type_embedding"]


def attention_probs(q: Tensor, k: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """softmax(QK^T/sqrt(dk)) of art:203-209 -> (B,H,Lq,Lk); what the dropout hook of eeg_metrics.py sees."""
    B, Lq, D = q.shape
    dk = D // heads
    qh = F.linear(q, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    kh = F.linear(k, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    return F.softmax(torch.matmul(qh, kh.transpose(-2, -1)) / math.sqrt(dk), dim=-1)


def mha(q: Tensor, k: Tensor, v: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """MultiHeadAttention.forward (art:202-213) without mask, eval mode."""
    B, _, D = q.shape
    dk = D // heads
    vh = F.linear(v, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    ctx = torch.matmul(attention_probs(q, k, sd, pre, heads), vh).transpose(1, 2).contiguous().view(B, -1, D)
    return F.linear(ctx, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def encoder(x: Tensor, sd: Dict[str, Tensor], cfg: EEGConfig, pre: str = "encoder.") -> Tensor:
    """TransformerEncoder (art:326-328) of post-LN blocks (art:292-296), eval mode."""
    D = cfg.d_model
    for i in range(cfg.num_layers):
        p = f"{pre}layers.{i}."
        h = mha(x, x, x, sd, p + "mha.", cfg.num_heads)
        x = F.layer_norm(x + h, (D,), sd[p + "ln1.weight"], sd[p + "ln1.bias"], 1e-5)
        h = F.linear(F.relu(F.linear(x, sd[p + "ffn.linear1.weight"], sd[p + "ffn.linear1.bias"])),
                     sd[p + "ffn.linear2.weight"], sd[p + "ffn.linear2.bias"])
        x = F.layer_norm(x + h, (D,), sd[p + "ln2.weight"], sd[p + "ln2.bias"], 1e-5)
    return F.layer_norm(x, (D,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)


def cross_brain(z1: Tensor, z2: Tensor, sd: Dict[str, Tensor], cfg: EEGConfig, pre: str = "cross_attn."):
    """CrossBrainAttention.forward (det:966-974): one MHA and one LayerNorm shared by both directions."""
    D = cfg.d_model
    w, b = sd[pre + "norm.weight"], sd[pre + "norm.bias"]
    o1 = F.layer_norm(z1 + mha(z1, z2, z2, sd, pre + "cross_attn.", cfg.num_heads), (D,), w, b, 1e-5)
    o2 = F.layer_norm(z2 + mha(z2, z1, z1, sd, pre + "cross_attn.", cfg.num_heads), (D,), w, b, 1e-5)
    return o1, o2


def symmetric_fusion(z1: Tensor, z2: Tensor, sd: Dict[str, Tensor], pre: str = "symmetric_fusion.") -> Tensor:
    """det:933-941."""
    return F.linear(torch.cat([z1 + z2, z1 * z2, (z1 - z2).abs()], dim=-1), sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def mlp_head(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """Linear-ReLU-(Dropout)-Linear heads (det:1074-1079, 1100-1105), eval mode."""
    return F.linear(F.relu(F.linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"])), sd[pre + "3.weight"], sd[pre + "3.bias"])


# ------------------------------------------------------------------------------------------------
# full forward
# ------------------------------------------------------------------------------------------------
def dual_eeg_forward(sd: Dict[str, Tensor], eeg1: Tensor, eeg2: Tensor, cfg: EEGConfig,
                     labels: Optional[Tensor] = None, ibs_matrices: Optional[Tensor] = None,
                     taps: Optional[dict] = None) -> Dict[str, Tensor]:
    """DualEEGTransformer.forward (det:1110-1253), eval mode (dropout off).

    ``ibs_matrices`` lets a caller pass pre-computed connectivity matrices (the generator has no
    parameters), which keeps repeated oracle calls on one fixture cheap.  ``taps`` (a dict) collects the tensors the
    reference's analysis hooks observe: 'spec_conv3' (list: player 1, player 2), 'cross_probs' (list: softmax of
    z1->z2, z2->z1, what the hook on cross_attn.cross_attn.dropout receives).
    """
    B = eeg1.shape[0]
    h1, h2 = temporal_conv(eeg1, sd, cfg), temporal_conv(eeg2, sd, cfg)
    ibs_tokens = None
    if cfg.use_ibs:
        if cfg.use_robust_ibs:
            if ibs_matrices is None:
                ibs_matrices = ibs_connectivity(eeg1, eeg2, cfg.sampling_rate, cfg.ibs_feature_type)
            ibs_tokens = ibs_tokenize(ibs_matrices, sd, cfg)
        else:
            f = ibs_scalar_features(eeg1, eeg2, cfg.sampling_rate)
            ibs_tokens = mlp_head(f, sd, "ibs_generator.proj.").unsqueeze(1)   # det:213-218, 464
    parts1 = [sd["cls_token"].expand(B, -1, -1)]
    parts2 = [sd["cls_token"].expand(B, -1, -1)]
    if ibs_tokens is not None:
        parts1.append(ibs_tokens)
        parts2.append(ibs_tokens)
    if cfg.use_spectrogram:
        parts1.append(spectrogram_tokens(eeg1, sd, cfg, taps=taps))
        parts2.append(spectrogram_tokens(eeg2, sd, cfg, taps=taps))
    parts1.append(h1)
    parts2.append(h2)
    s1, s2 = torch.cat(parts1, 1), torch.cat(parts2, 1)
    L = s1.shape[1]
    if L > sd["pos_embed.pos_embed.weight"].shape[0]:
        raise IndexError("sequence longer than max_len")          # art:122-123 embedding lookup
    pos = sd["pos_embed.pos_embed.weight"][:L]
    z1, z2 = encoder(s1 + pos, sd, cfg), encoder(s2 + pos, sd, cfg)
    if cfg.use_cross_attention:
        if taps is not None:
            with torch.no_grad():
                taps["cross_probs"] = [attention_probs(z1, z2, sd, "cross_attn.cross_attn.", cfg.num_heads),
                                       attention_probs(z2, z1, sd, "cross_attn.cross_attn.", cfg.num_heads)]
        z1, z2 = cross_brain(z1, z2, sd, cfg)
    cls1, cls2 = z1[:, 0], z2[:, 0]
    offset = 1 + cfg.num_ibs_tokens + (cfg.in_channels if cfg.use_spectrogram else 0)   # det:1198-1202
    mp1, mp2 = z1[:, offset:].mean(1), z2[:, offset:].mean(1)
    logits = mlp_head(torch.cat([symmetric_fusion(cls1, cls2, sd), mp1, mp2], -1), sd, "classifier.")
    out = {"logits": logits, "cls1": cls1, "cls2": cls2}
    if cfg.use_ibs:
        tok = z1[:, 1:1 + cfg.num_ibs_tokens].mean(1) if cfg.use_robust_ibs else z1[:, 1]   # det:1219-1230
        out["ibs_logits"] = mlp_head(tok, sd, "ibs_classifier.")
        out["ibs_token"] = tok
    if labels is not None:
        out["loss"] = out["loss_ce"] = F.cross_entropy(logits, labels)
        if cfg.use_ibs:
            out["loss_ibs_cls"] = F.cross_entropy(out["ibs_logits"], labels)
    return out


# ------------------------------------------------------------------------------------------------
# auxiliary losses (det:1255-1371)
# ------------------------------------------------------------------------------------------------
def symmetry_loss(cls1: Tensor, cls2: Tensor) -> Tensor:
    return F.mse_loss(cls1, cls2)


def ibs_alignment_loss(ibs_token: Tensor, cls1: Tensor, cls2: Tensor, temperature: float = 0.07) -> Tensor:
    n = F.normalize(ibs_token, dim=-1)
    allc = torch.cat([F.normalize(cls1, dim=-1), F.normalize(cls2, dim=-1)], 0)
    return F.cross_entropy(n @ allc.T / temperature, torch.arange(ibs_token.shape[0], device=ibs_token.device))


def ibs_contrastive_loss(tokens: Tensor, labels: Tensor, temperature: float = 0.07) -> Tensor:
    B = tokens.shape[0]
    t = F.normalize(tokens, p=2, dim=1)
    sim = t @ t.t() / temperature
    eye = torch.eye(B, dtype=torch.bool, device=tokens.device)
    pos = (labels[:, None] == labels[None, :]).float().masked_fill(eye, 0)
    has = pos.sum(1) > 0
    if has.sum() == 0:
        return torch.tensor(0.0, device=tokens.device)
    e = torch.exp(sim)
    loss = -torch.log((e * pos).sum(1) / (e.masked_fill(eye, 0).sum(1) + 1e-8) + 1e-8)
    return loss[has].mean()


# ------------------------------------------------------------------------------------------------
# parameter construction (shapes of Appendix C) for oracle-vs-CUDA tests at sizes with no fixture
# ------------------------------------------------------------------------------------------------
def init_state_dict(cfg: EEGConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Random parameters with the reference's shapes and roughly its init scales (not its RNG stream)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    D, C = cfg.d_model, cfg.in_channels

    def lin(name, out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    def conv(name, out_c, in_c, *k):
        fan = in_c * int(torch.tensor(k).prod())
        bound = 1.0 / math.sqrt(fan)
        sd[name + ".weight"] = (torch.rand(out_c, in_c, *k, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(out_c, generator=g) * 2 - 1) * bound

    def ln(name, n):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(n, generator=g)

    sd["cls_token"] = torch.randn(1, 1, D, generator=g)
    conv("temporal_conv.convs.0", D, C, cfg.conv_kernel_size)
    for i in range(1, cfg.conv_layers):
        conv(f"temporal_conv.convs.{i}", D, D, cfg.conv_kernel_size)
    if cfg.use_spectrogram:
        sd["spectrogram_generator.window"] = torch.hann_window(cfg.spec_n_fft)
        conv("spectrogram_generator.spec_conv.0", 32, 1, 3, 3)
        conv("spectrogram_generator.spec_conv.3", 64, 32, 3, 3)
        lin("spectrogram_generator.proj.0", 2 * D, 1024)
        lin("spectrogram_generator.proj.3", D, 2 * D)
    if cfg.use_ibs:
        if cfg.use_robust_ibs:
            nt = cfg.num_ibs_tokens
            sd["ibs_tokenizer.type_embedding"] = 0.02 * torch.randn(1, nt, D, generator=g)
            if cfg.ibs_instance_norm:
                ln("ibs_tokenizer.instance_norm", C * C)
            lin("ibs_tokenizer.bottleneck.0", 64, C * C)
            lin("ibs_tokenizer.bottleneck.3", D, 64)
        else:
            lin("ibs_generator.proj.0", 2 * D, 28)
            lin("ibs_generator.proj.3", D, 2 * D)
        lin("ibs_classifier.0", D // 2, D)
        lin("ibs_classifier.3", cfg.num_classes, D // 2)
    sd["pos_embed.pos_embed.weight"] = torch.randn(cfg.max_len, D, generator=g)
    for i in range(cfg.num_layers):
        p = f"encoder.layers.{i}."
        for n in ("q", "k", "v", "out"):
            lin(p + f"mha.{n}_proj", D, D)
        ln(p + "ln1", D)
        lin(p + "ffn.linear1", cfg.d_ff, D)
        lin(p + "ffn.linear2", D, cfg.d_ff)
        ln(p + "ln2", D)
    ln("encoder.norm", D)
    if cfg.use_cross_attention:
        for n in ("q", "k", "v", "out"):
            lin(f"cross_attn.cross_attn.{n}_proj", D, D)
        ln("cross_attn.norm", D)
    lin("symmetric_fusion.proj", D, 3 * D)
    lin("classifier.0", D, 3 * D)
    lin("classifier.3", cfg.num_classes, D)
    return sd
