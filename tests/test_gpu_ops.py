"""GPU parity tests of the individual fused ops (through the C ABI) against plain fp32 PyTorch-CPU restatements
of the reference op chains (oracle/).  Run on the B200 box:  pytest -m gpu"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_state_dict, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200 import _lib as L
    from eyegaze_multimodal_b200 import ops
    from eyegaze_multimodal_b200.precision import precision
from oracle import eeg as O
from oracle import fuzzy as FZ

DEV = "cuda:0"
TOL = {"fp32": dict(atol=2e-5, rtol=2e-5), "bf16": dict(atol=6e-2, rtol=6e-2)}


def _param(t):
    return torch.nn.Parameter(t.clone().to(DEV))


def _close(got, want, mode, scale=1.0, msg=""):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    tol = TOL[mode]
    denom = want.abs().max().item() + 1e-12
    err = (got - want).abs().max().item()
    assert err <= tol["atol"] * scale + tol["rtol"] * denom, f"{msg}: max abs err {err:.3e} (ref max {denom:.3e})"


def _dt(mode):
    return torch.float32 if mode == "fp32" else torch.bfloat16


def _code(mode):
    return L.F32 if mode == "fp32" else L.BF16


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
# (1000, 72, 64) and (2100, 200, 96): N is a multiple of 8 but not of 32 -- the last 32-column chunk of the row-layout
# epilogue is partial and the TMA store clips it; (20000, 512, 64): enough 256-row tiles for the CTA-pair kernel
@pytest.mark.parametrize("shape", [(70, 96, 64), (300, 256, 128), (1000, 768, 256), (1000, 72, 64), (2100, 200, 96),
                                   (20000, 512, 64)])
def test_linear_fwd_bwd(cuda_device, mode, shape):
    M, N, K = shape
    torch.manual_seed(0)
    x = torch.randn(M, K) * 0.5
    w = torch.randn(N, K) / math.sqrt(K)
    b = torch.randn(N) * 0.1
    res = torch.randn(M, N) * 0.3
    gy = torch.randn(M, N)
    xr, wr, br, rr = [t.clone().requires_grad_(True) for t in (x, w, b, res)]
    q = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)
    # (a) bias + residual (out_proj + residual of art.py:293), (b) bias + relu (conv / head layers)
    for variant in ("residual", "relu"):
        for t in (xr, wr, br, rr):
            t.grad = None
        lin = F.linear(q(xr), q(wr), br)
        yr = lin + q(rr) if variant == "residual" else F.relu(lin)
        yr.backward(gy)
        xg = x.to(DEV).to(_dt(mode)).requires_grad_(True)
        rg = res.to(DEV).to(_dt(mode)).requires_grad_(True)
        wg, bg = _param(w), _param(b)
        if variant == "residual":
            y = ops.linear(xg, wg, bg, residual=rg)
        else:
            y = ops.linear(xg, wg, bg, act=L.ACT_RELU)
        assert y.dtype == _dt(mode)
        y.backward(gy.to(DEV).to(_dt(mode)))
        _close(y, yr, mode, msg="y")
        _close(xg.grad, xr.grad, mode, msg="dx")
        if variant == "residual":
            _close(rg.grad, rr.grad, mode, msg="dres")
        _close(wg.grad, wr.grad, mode, scale=math.sqrt(M), msg="dW")
        _close(bg.grad, br.grad, mode, scale=math.sqrt(M), msg="db")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("act", ["relu", "gelu"])
def test_mlp2_fwd_bwd(cuda_device, mode, act):
    torch.manual_seed(1)
    G, R, K, Hd, N = 6, 37, 64, 128, 64
    x = torch.randn(G, R, K)
    w1, b1 = torch.randn(Hd, K) / 8, torch.randn(Hd) * 0.1
    w2, b2 = torch.randn(N, Hd) / 11, torch.randn(N) * 0.1
    gy = torch.randn(G, R, N)
    ps = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    fn = F.relu if act == "relu" else F.gelu
    q = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)
    h = fn(F.linear(q(ps[0]), q(ps[1]), ps[2]))
    yr = F.linear(q(h), q(ps[3]), ps[4]) + q(ps[0])
    yr.backward(gy)
    xg = x.to(DEV).to(_dt(mode)).requires_grad_(True)
    prm = [_param(t) for t in (w1, b1, w2, b2)]
    y = ops.mlp2(xg, *prm, L.ACT_RELU if act == "relu" else L.ACT_GELU, residual=xg)
    y.backward(gy.to(DEV).to(_dt(mode)))
    _close(y, yr, mode, msg="y")
    _close(xg.grad, ps[0].grad, mode, msg="dx")
    for g, r, nm in zip(prm, ps[1:], ["dw1", "db1", "dw2", "db2"]):
        _close(g.grad, r.grad, mode, scale=math.sqrt(G * R), msg=nm)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("D", [16, 256, 768])
def test_layernorm(cuda_device, mode, D):
    torch.manual_seed(2)
    x = torch.randn(5, 33, D) * 2 + 0.5
    g, b = 1 + 0.1 * torch.randn(D), 0.1 * torch.randn(D)
    gy = torch.randn(5, 33, D)
    xr, gr, br = [t.clone().requires_grad_(True) for t in (x, g, b)]
    xq = xr.bfloat16().float() if mode == "bf16" else xr
    yr = F.layer_norm(xq, (D,), gr, br, 1e-5)
    yr.backward(gy)
    xg = x.to(DEV).to(_dt(mode)).requires_grad_(True)
    gg, bg = _param(g), _param(b)
    y = ops.layernorm(xg, gg, bg, 1e-5)
    y.backward(gy.to(DEV).to(_dt(mode)))
    _close(y, yr, mode, msg="y")
    _close(xg.grad, xr.grad, mode, msg="dx")
    _close(gg.grad, gr.grad, mode, scale=10, msg="dgamma")
    _close(bg.grad, br.grad, mode, scale=10, msg="dbeta")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("D", [256, 768])
def test_layernorm_residual_block(cuda_device, mode, D):
    """Pre-norm residual block y = x + W LN(x): the gradient that reaches x through the residual operand handed back
    by ops.layernorm_residual is added inside the LayerNorm backward kernel (egb_layernorm_bwd_res)."""
    torch.manual_seed(12)
    x = torch.randn(7, 29, D) * 1.5
    g, b = 1 + 0.1 * torch.randn(D), 0.1 * torch.randn(D)
    w, wb = torch.randn(D, D) / math.sqrt(D), 0.1 * torch.randn(D)
    gy = torch.randn(7, 29, D)
    xr, gr, br, wr, wbr = [t.clone().requires_grad_(True) for t in (x, g, b, w, wb)]
    q = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)
    xq = q(xr)
    yr = F.linear(q(F.layer_norm(xq, (D,), gr, br, 1e-6)), q(wr), wbr) + xq
    yr.backward(gy)
    xg = x.to(DEV).to(_dt(mode)).requires_grad_(True)
    gg, bg, wg, wbg = _param(g), _param(b), _param(w), _param(wb)
    h, xres = ops.layernorm_residual(xg, gg, bg, 1e-6)
    y = ops.linear(h, wg, wbg, residual=xres)
    y.backward(gy.to(DEV).to(_dt(mode)))
    _close(y, yr, mode, msg="y")
    _close(xg.grad, xr.grad, mode, msg="dx (LayerNorm path + residual path)")
    _close(gg.grad, gr.grad, mode, scale=10, msg="dgamma")
    _close(bg.grad, br.grad, mode, scale=10, msg="dbeta")
    _close(wg.grad, wr.grad, mode, scale=math.sqrt(7 * 29), msg="dw")


@pytest.mark.parametrize("act_bwd", ["mul", "relu"])
def test_gemm_epilogue_column_sums(cuda_device, act_bwd):
    """egb_gemm_desc.c_colsum: the tensor-core epilogue accumulates the column sums of the tile it stores (bias gradient
    of the layer whose pre-activation gradient the GEMM produces); compared with a separate pass over the output."""
    torch.manual_seed(13)
    M, N, K = 1000, 512, 256                     # dpre[M, N] . W[N, K] -> dx[M, K]
    dpre = (torch.randn(M, N, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=DEV) / math.sqrt(N)).bfloat16()
    aux = torch.randn(M, K, device=DEV).bfloat16()
    if act_bwd == "relu":
        aux = torch.relu(aux)
    code = L.ACTBWD_MUL if act_bwd == "mul" else L.ACTBWD_RELU_MASK
    fused = ops.small_zeros((K,), torch.device(DEV))
    dx1 = ops._grad_input(dpre, w, (M, K), L.BF16, act_bwd=code, aux=aux, colsum_out=fused)
    dx0 = ops._grad_input(dpre, w, (M, K), L.BF16, act_bwd=code, aux=aux)
    assert torch.equal(dx0, dx1)
    want = dx0.float().sum(0)
    sep = ops.colsum(dx0, K)
    ref_scale = dx0.float().abs().sum(0).max().item()
    assert (sep - want).abs().max().item() <= 1e-3 * ref_scale
    # the fused sums are taken before the bf16 rounding of the stored values: agree to bf16 accuracy of the column mass
    assert (fused - want).abs().max().item() <= 4e-3 * ref_scale


def _ref_attention(q, k, v, H, kv_shift=0):
    S, Lq, D = q.shape
    dk = D // H
    if kv_shift:
        k, v = torch.roll(k, -kv_shift, 0), torch.roll(v, -kv_shift, 0)
    qh = q.view(S, Lq, H, dk).transpose(1, 2)
    kh = k.view(S, -1, H, dk).transpose(1, 2)
    vh = v.view(S, -1, H, dk).transpose(1, 2)
    a = F.softmax(qh @ kh.transpose(-2, -1) / math.sqrt(dk), dim=-1)
    return (a @ vh).transpose(1, 2).reshape(S, Lq, D), a


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [(4, 33, 33, 64, 4, 0), (4, 139, 139, 256, 8, 2), (2, 197, 197, 384, 6, 0)])
def test_attention_packed(cuda_device, mode, cfg):
    S, Lq, Lk, D, H, shift = cfg
    torch.manual_seed(3)
    qkv = torch.randn(S, Lq, 3 * D) * 0.7
    go = torch.randn(S, Lq, D)
    r = qkv.clone().requires_grad_(True)
    rq = r.bfloat16().float() if mode == "bf16" else r
    o_ref, probs_ref = _ref_attention(rq[..., :D], rq[..., D:2 * D], rq[..., 2 * D:], H, shift)
    o_ref.backward(go)
    g = qkv.to(DEV).to(_dt(mode)).requires_grad_(True)
    o, probs = ops.attention_packed(g, H, kv_shift=shift, want_probs=True)
    o.backward(go.to(DEV).to(_dt(mode)))
    _close(o, o_ref, mode, msg="ctx")
    _close(probs, probs_ref, mode, msg="probs")
    _close(g.grad, r.grad, mode, msg="dqkv")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_attention_unequal_lengths(cuda_device, mode):
    """BASELINE config 3: 64 gaze-side tokens x 128 EEG-side tokens, d=256, 8 heads."""
    torch.manual_seed(4)
    S, Lq, Lk, D, H = 3, 64, 128, 256, 8
    q, k, v = torch.randn(S, Lq, D), torch.randn(S, Lk, D), torch.randn(S, Lk, D)
    go = torch.randn(S, Lq, D)
    rs = [t.clone().requires_grad_(True) for t in (q, k, v)]
    rq = [t.bfloat16().float() if mode == "bf16" else t for t in rs]
    o_ref, _ = _ref_attention(*rq, H)
    o_ref.backward(go)
    gs = [t.to(DEV).to(_dt(mode)).requires_grad_(True) for t in (q, k, v)]
    o = ops.attention(*gs, H)
    o.backward(go.to(DEV).to(_dt(mode)))
    _close(o, o_ref, mode, msg="ctx")
    for a, b, nm in zip(gs, rs, "qkv"):
        _close(a.grad, b.grad, mode, msg="d" + nm)


def test_attention_dropout_statistics(cuda_device):
    torch.manual_seed(5)
    S, Lq, D, H = 8, 64, 64, 4
    qkv = torch.randn(S, Lq, 3 * D, device=DEV)
    base = ops.attention_packed(qkv, H)
    outs = torch.stack([ops.attention_packed(qkv, H, p=0.3) for _ in range(64)])
    assert (outs[0] - outs[1]).abs().max() > 1e-3          # masks differ between calls
    assert (outs.mean(0) - base).abs().mean() < 0.05       # inverted dropout is unbiased


@pytest.mark.parametrize("cfg", [(6, 139, 139, 256, 8, 3), (2, 197, 197, 384, 6, 0), (2, 235, 235, 256, 8, 1),
                                 (3, 256, 256, 128, 2, 0), (2, 17, 40, 64, 2, 0)])
def test_attention_tensor_core_vs_fp32_kernel_same_dropout_mask(cuda_device, cfg):
    """The tcgen05 kernels (bf16) and the CUDA-core kernels (fp32) draw the same counter-based dropout mask for the
    same seed, so outputs and gradients must agree to bf16 accuracy -- forward, dQ, dK, dV, with kv_shift."""
    S, Lq, Lk, D, H, shift = cfg
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(S, Lq, D, generator=g) * 0.8).to(DEV)
    k = (torch.randn(S, Lk, D, generator=g) * 0.8).to(DEV)
    v = torch.randn(S, Lk, D, generator=g).to(DEV)
    go = torch.randn(S, Lq, D, generator=g).to(DEV)
    res = {}
    for mode in ("fp32", "bf16"):
        torch.manual_seed(1000 + len(res))      # changing the base seed ...
        ops.next_seed()
        torch.manual_seed(77)                   # ... and restoring it restarts the op-seed sequence
        ts = [t.detach().clone().to(_dt(mode)).requires_grad_(True) for t in (q, k, v)]
        o = ops.AttentionFn.apply(*ts, H, shift, 0.25, 1.0 / math.sqrt(D // H), False, False)
        o.backward(go.to(_dt(mode)))
        res[mode] = [o.detach().float()] + [t.grad.float() for t in ts]
    for a, b, nm in zip(res["bf16"], res["fp32"], ["o", "dq", "dk", "dv"]):
        err = (a - b).abs().max().item()
        assert err <= 3e-2 * b.abs().max().item() + 1e-3, f"{nm}: {err:.3e} vs max {b.abs().max().item():.3e}"


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [(3, 8, 256, 32), (2, 32, 1024, 256), (2, 62, 512, 64)])
def test_temporal_conv(cuda_device, mode, cfg):
    B, C, T, D = cfg
    torch.manual_seed(6)
    e1, e2 = torch.randn(B, C, T), torch.randn(B, C, T)
    w1, b1 = torch.randn(D, C, 25) / math.sqrt(25 * C), torch.randn(D) * 0.1
    w2, b2 = torch.randn(D, D, 25) / math.sqrt(25 * D), torch.randn(D) * 0.1
    ps = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    q = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)

    def ref(x):
        h = F.relu(F.conv1d(q(x), q(ps[0]), ps[1], stride=4, padding=12))
        h = F.relu(F.conv1d(q(h), q(ps[2]), ps[3], stride=4, padding=12))
        return h.permute(0, 2, 1)
    hr = torch.cat([ref(e1), ref(e2)], 0)
    gh = torch.randn_like(hr)
    hr.backward(gh)
    prm = [_param(t) for t in (w1, b1, w2, b2)]
    h = ops.temporal_conv(e1.to(DEV), e2.to(DEV), [prm[0], prm[2]], [prm[1], prm[3]], _code(mode), 4, 0.0)
    assert h.shape == hr.shape
    h.backward(gh.to(DEV).to(_dt(mode)))
    _close(h, hr, mode, msg="h")
    for g, r, nm in zip(prm, ps, ["dw1", "db1", "dw2", "db2"]):
        _close(g.grad, r.grad, mode, scale=math.sqrt(B * T / 4), msg=nm)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [(2, 4, 256), (2, 8, 1024)])
def test_spectrogram_cnn(cuda_device, mode, cfg):
    B, C, T = cfg
    torch.manual_seed(7)
    e1, e2 = torch.randn(B, C, T), torch.randn(B, C, T)
    ocfg = O.EEGConfig(in_channels=C)
    win = torch.hann_window(128)
    w1, b1 = torch.randn(32, 1, 3, 3) / 3, torch.randn(32) * 0.1
    w2, b2 = torch.randn(64, 32, 3, 3) / 17, torch.randn(64) * 0.1
    ps = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    q = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)

    def ref(x):
        img = O.spectrogram_logmag(x, win, ocfg)
        h = F.max_pool2d(F.relu(F.conv2d(img, ps[0], ps[1], padding=1)), 2)
        h = F.relu(F.conv2d(q(h), q(ps[2]), ps[3], padding=1))
        return F.adaptive_avg_pool2d(h, (4, 4)).flatten(1)
    fr = torch.cat([ref(e1), ref(e2)], 0)
    gf = torch.randn_like(fr)
    fr.backward(gf)
    prm = [_param(t) for t in (w1, b1, w2, b2)]
    f = ops.spectrogram_cnn(e1.to(DEV), e2.to(DEV), win.to(DEV), *prm, _code(mode), 128, 64, 64)
    f.backward(gf.to(DEV).to(_dt(mode)))
    _close(f, fr, mode, msg="pooled")
    for g, r, nm in zip(prm, ps, ["dw1", "db1", "dw2", "db2"]):
        _close(g.grad, r.grad, mode, scale=math.sqrt(B * C * 64), msg=nm)


def _ibs_compare(got, ref, T):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    for f in (0, 3, 4, 5, 6):
        d = np.abs(got[:, :, f] - ref[:, :, f]).max()
        assert d <= 3e-5, f"feature {f}: max abs err {d}"
    dpli = np.abs(got[:, :, 1] - ref[:, :, 1])
    assert (dpli > 1e-5).mean() <= 0.05 and (dpli <= 8.0 / T + 1e-6).all(), "PLI differs by more than a few sign flips"
    dw = np.abs(got[:, :, 2] - ref[:, :, 2])
    assert (dw > 1e-4).mean() <= 0.05 and dw.max() < 0.05


@pytest.mark.parametrize("name", ["ibs_small.npz", "ibs_c32.npz"])
def test_ibs_connectivity_golden(cuda_device, name):
    """CUDA kernels vs matrices produced by the UNMODIFIED reference (flip-aware metric)."""
    g = load_golden(name)
    e1, e2 = torch.from_numpy(g["eeg1"]).to(DEV), torch.from_numpy(g["eeg2"]).to(DEV)
    got = ops.ibs_connectivity(e1, e2, 256.0, O.IBS_BANDS, list(range(7)))
    _ibs_compare(got.cpu().numpy(), g["matrices"], e1.shape[-1])
    sub = ops.ibs_connectivity(e1, e2, 256.0, O.IBS_BANDS, [3, 4, 6])
    assert torch.equal(sub, got[:, :, [3, 4, 6]])


def test_ibs_connectivity_vs_oracle_c64(cuda_device):
    """BASELINE config 5 geometry (64 channels x 2048 samples), one trial, against the CPU oracle."""
    from eyegaze_multimodal_b200.synth import eeg_pair_batch
    e1, e2 = eeg_pair_batch(1, 64, 2048, seed=5)
    want = O.ibs_connectivity(e1, e2, 256.0, "all", chunk=1)
    got = ops.ibs_connectivity(e1.to(DEV), e2.to(DEV), 256.0, O.IBS_BANDS, list(range(7)))
    _ibs_compare(got.cpu().numpy(), want.numpy(), 2048)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_instnorm_tokens(cuda_device, mode):
    torch.manual_seed(8)
    x = torch.rand(3, 42, 64)
    g, b = 1 + 0.1 * torch.randn(64), 0.1 * torch.randn(64)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.instance_norm(x.permute(0, 2, 1), weight=gr, bias=br, eps=1e-5).permute(0, 2, 1)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    gg, bg = _param(g), _param(b)
    y = ops.instnorm_tokens(x.to(DEV), gg, bg, _code(mode), True)
    y.backward(gy.to(DEV).to(_dt(mode)))
    _close(y, yr, mode, msg="y")
    _close(gg.grad, gr.grad, mode, scale=10, msg="dgamma")
    _close(bg.grad, br.grad, mode, scale=10, msg="dbeta")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_seq_assemble_and_tail(cuda_device, mode):
    torch.manual_seed(9)
    B, D, n_ibs, C, n_h = 3, 32, 5, 4, 6
    Lq = 1 + n_ibs + C + n_h
    cls, pos = torch.randn(1, 1, D), torch.randn(Lq + 3, D)
    ibs, spec, h = torch.randn(B, n_ibs, D), torch.randn(2 * B, C, D), torch.randn(2 * B, n_h, D)
    rs = [t.clone().requires_grad_(True) for t in (cls, pos, ibs, spec, h)]
    q = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)
    seq = torch.cat([rs[0].expand(2 * B, -1, -1), q(rs[2]).repeat(2, 1, 1), q(rs[3]), q(rs[4])], 1) + rs[1][:Lq]
    z = q(seq) if mode == "bf16" else seq
    off = 1 + n_ibs + C
    c1, c2 = z[:B, 0], z[B:, 0]
    sym = torch.cat([c1 + c2, c1 * c2, (c1 - c2).abs()], -1)
    mp = torch.cat([z[:B, off:].mean(1), z[B:, off:].mean(1)], -1)
    ip = z[:B, 1:1 + n_ibs].mean(1)
    ws = [torch.randn_like(t) for t in (c1, c2, sym, mp, ip)]
    sum((a * w).sum() for a, w in zip((c1, c2, sym, mp, ip), ws)).backward()
    cg, pg = _param(cls), _param(pos)
    gs = [t.to(DEV).to(_dt(mode)).requires_grad_(True) for t in (ibs, spec, h)]
    x = ops.seq_assemble(cg, pg, gs[0], gs[1], gs[2], _code(mode))
    _close(x, seq, mode, msg="seq")
    outs = ops.tail_pool(x, n_ibs, off, False)
    for o, r, nm in zip(outs, (c1, c2, sym, mp, ip), ["cls1", "cls2", "sym", "mp", "ibs_pool"]):
        _close(o, r, mode, msg=nm)
    sum((a * w.to(DEV)).sum() for a, w in zip(outs, ws)).backward()
    _close(cg.grad, rs[0].grad, mode, scale=3, msg="dcls")
    _close(pg.grad, rs[1].grad, mode, scale=3, msg="dpos")
    for g, r, nm in zip(gs, rs[2:], ["dibs", "dspec", "dh"]):
        _close(g.grad, r.grad, mode, scale=3, msg=nm)


def test_cross_entropy(cuda_device):
    torch.manual_seed(10)
    x = torch.randn(37, 3)
    y = torch.randint(0, 3, (37,))
    xr = x.clone().requires_grad_(True)
    lr = F.cross_entropy(xr, y)
    (lr * 1.7).backward()
    xg = x.to(DEV).requires_grad_(True)
    lg = ops.cross_entropy(xg, y.to(DEV))
    (lg * 1.7).backward()
    assert abs(lg.item() - lr.item()) < 1e-6
    assert (xg.grad.cpu() - xr.grad).abs().max() < 1e-7


@pytest.mark.parametrize("mode_name", ["full", "no_temperature", "no_fuzzification", "fixed_weights"])
def test_fuzzy_gating_golden(cuda_device, mode_name):
    """Fused fuzzy-fusion kernel vs outputs / gradients of the UNMODIFIED reference module."""
    g = load_golden("fuzzy_fusion.npz")
    names = ["tau_img", "tau_eeg", "c_reliable", "c_unreliable_img", "c_unreliable_eeg", "log_sigma_reliable_img",
             "log_sigma_reliable_eeg", "log_sigma_unreliable_img", "log_sigma_unreliable_eeg", "beta"]
    init = FZ.init_params()
    params = [init[n].clone().to(DEV).requires_grad_(n != "c_reliable") for n in names]
    img = torch.from_numpy(g["img"]).to(DEV).requires_grad_(True)
    eeg = torch.from_numpy(g["eeg"]).to(DEV).requires_grad_(True)
    fused, alpha, aux = ops.fuzzy_gating(img, eeg, ops.FUZZY_MODES[mode_name], 0.1, 1e-8, 1e-8, params)
    np.testing.assert_allclose(fused.detach().cpu().numpy(), g[f"{mode_name}::fused"], atol=5e-6)
    np.testing.assert_allclose(alpha.detach().cpu().numpy(), g[f"{mode_name}::alpha"], atol=5e-6)
    np.testing.assert_allclose(aux[:-1, 0].cpu().numpy(), g[f"{mode_name}::H_img"], atol=2e-6)
    (fused * torch.arange(1, 4, device=DEV)).sum().backward()
    np.testing.assert_allclose(img.grad.cpu().numpy(), g[f"{mode_name}::grad_img"], atol=1e-5, rtol=2e-5)
    np.testing.assert_allclose(eeg.grad.cpu().numpy(), g[f"{mode_name}::grad_eeg"], atol=1e-5, rtol=2e-5)
    for n, p in zip(names, params):
        key = f"{mode_name}::grad::{n}"
        if key in g:
            got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(g[key])
            np.testing.assert_allclose(got, g[key], atol=2e-5, rtol=2e-5, err_msg=n)


def test_library_has_no_cpu_path(cuda_device):
    with pytest.raises(RuntimeError):
        ops.linear(torch.randn(4, 8), torch.nn.Parameter(torch.randn(8, 8)))


# The fp32 heads of a step are "small" problems (M = batch rows): their K loop is split over a thread-block cluster
# (non-accumulating launches, DSMEM reduction) or over atomically accumulating CTAs (weight gradients).
# Under precision("bf16") (the throughput mode, whose heads stay fp32) the cluster split is on; under precision("fp32")
# (parity mode) ops.gemm asks for the in-order K loop.
@pytest.mark.parametrize("dims", [(256, 1024, 512, 256), (256, 768, 256, 3), (100, 300, 200, 40), (256, 128, 128, 64)])
@pytest.mark.parametrize("act", ["relu", "gelu"])
@pytest.mark.parametrize("policy", ["fp32", "bf16"])
def test_mlp2_small_rows_fp32_split_k(cuda_device, dims, act, policy):
    R, K, Hd, N = dims
    torch.manual_seed(3)
    x = torch.randn(R, K)
    w1, b1 = torch.randn(Hd, K) / math.sqrt(K), torch.randn(Hd) * 0.1
    w2, b2 = torch.randn(N, Hd) / math.sqrt(Hd), torch.randn(N) * 0.1
    gy = torch.randn(R, N)
    ps = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    fn = F.relu if act == "relu" else F.gelu
    yr = F.linear(fn(F.linear(ps[0], ps[1], ps[2])), ps[3], ps[4])
    yr.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    prm = [_param(t) for t in (w1, b1, w2, b2)]
    with precision(policy):
        y = ops.mlp2(xg, *prm, L.ACT_RELU if act == "relu" else L.ACT_GELU)
        y.backward(gy.to(DEV))
    assert y.dtype == torch.float32
    _close(y, yr, "fp32", msg="y")
    _close(xg.grad, ps[0].grad, "fp32", msg="dx")
    for g, r, nm in zip(prm, ps[1:], ["dw1", "db1", "dw2", "db2"]):
        _close(g.grad, r.grad, "fp32", msg=nm)
