"""GPU parity of the direct 3x3 convolution kernels (csrc/conv3x3.cu: forward 32 -> 64 + bias, data gradient 64 -> 32,
weight gradient) through their C-ABI entry points, against torch's fp32 conv2d on the same bf16-rounded inputs
(spec_conv[3] of the spectrogram CNN, dual_eeg_transformer.py:81-86).  Geometries cover the narrowest and the widest
padded rows the kernels accept, a single partial tile and ragged last tiles.  Run on the B200 box:  pytest -m gpu"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from eyegaze_multimodal_b200 import torch_ops as TO

DEV = "cuda:0"
# (images, H1, W1): Wp = W1 + 2 is the row shift of one kernel row
GEOMETRIES = [(1, 2, 3), (3, 5, 1), (5, 17, 20), (40, 9, 11), (2, 7, 60)]


def _q(t):
    return t.bfloat16().float()


def _padded(x):
    """(N, C, H1, W1) fp32 -> flat zero-bordered channels-last bf16 buffer [(N * RP + slack), C] on the device."""
    N, C, H1, W1 = x.shape
    Wp, Hp = W1 + 2, H1 + 2
    buf = torch.zeros(N, Hp, Wp, C)
    buf[:, 1:-1, 1:-1, :] = x.permute(0, 2, 3, 1)
    flat = torch.cat([buf.reshape(N * Hp * Wp, C), torch.zeros(2 * Wp + 8, C)])
    return flat.bfloat16().to(DEV).contiguous()


def _interior(buf, N, C, H1, W1):
    Wp, Hp = W1 + 2, H1 + 2
    img = buf[: N * Hp * Wp].float().cpu().reshape(N, Hp, Wp, C)
    return img[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)


def _case(geom, seed):
    N, H1, W1 = geom
    g = torch.Generator().manual_seed(seed)
    x = _q(torch.randn(N, 32, H1, W1, generator=g))
    w = _q(torch.randn(64, 32, 3, 3, generator=g) / 17.0)
    b = torch.randn(64, generator=g) * 0.1
    dy = _q(torch.randn(N, 64, H1, W1, generator=g))
    return x, w, b, dy


def _assert_close(got, want, rel, what):
    err = (got - want).abs().max().item()
    ref = want.abs().max().item() + 1e-12
    assert err <= rel * ref, "%s: max abs err %.3e against max |ref| %.3e" % (what, err, ref)


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_conv3x3_forward_matches_conv2d(cuda_device, geom):
    N, H1, W1 = geom
    Wp, RP = W1 + 2, (H1 + 2) * (W1 + 2)
    x, w, b, _ = _case(geom, 1)
    xp = _padded(x)
    wseg = torch.zeros(64, 3, 4, 32)
    wseg[:, :, :3, :] = w.permute(0, 2, 3, 1)                      # [o, kh, kw, c], kw = 3 slot unused
    wseg = wseg.reshape(64, 384).bfloat16().to(DEV)
    y = torch.full(((N * RP + 2 * Wp + 8) * 64,), float("nan"), dtype=torch.bfloat16, device=DEV)
    TO.call("conv3x3_c32_c64", xp, xp.shape[0], wseg, b.to(DEV), y, N * RP, Wp + 1, Wp)
    torch.cuda.synchronize()
    rows = y.view(-1, 64).float().cpu()
    assert torch.isnan(rows[: Wp + 1]).all() and torch.isnan(rows[N * RP + Wp + 1:]).all(), "rows outside [shift, shift + M) written"
    got = _interior(y.view(-1, 64), N, 64, H1, W1)
    want = F.conv2d(x, w, b, padding=1)
    _assert_close(got, want, 6e-3, "forward")                     # bf16 rounding of the stored output: 2^-9 relative


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_conv3x3_data_gradient_matches_autograd(cuda_device, geom):
    N, H1, W1 = geom
    Wp, RP = W1 + 2, (H1 + 2) * (W1 + 2)
    x, w, b, dy = _case(geom, 2)
    dyp = _padded(dy)
    wflip = torch.zeros(32, 3, 4, 64)
    wflip[:, :, :3, :] = w.flip(2, 3).permute(1, 2, 3, 0)          # [c, a, b, o] = W[o, c, 2 - a, 2 - b]
    wflip = wflip.reshape(32, 768).bfloat16().to(DEV)
    dx = torch.full(((N * RP + 2 * Wp + 8) * 32,), float("nan"), dtype=torch.bfloat16, device=DEV)
    TO.call("conv3x3_c64_c32", dyp, dyp.shape[0], wflip, dx, N * RP, Wp + 1, Wp)
    torch.cuda.synchronize()
    rows = dx.view(-1, 32).float().cpu()
    assert torch.isnan(rows[: Wp + 1]).all() and torch.isnan(rows[N * RP + Wp + 1:]).all(), "rows outside [shift, shift + M) written"
    got = _interior(dx.view(-1, 32), N, 32, H1, W1)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, w, b, padding=1).backward(dy)
    _assert_close(got, xr.grad, 6e-3, "data gradient")


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_conv3x3_weight_gradient_matches_autograd(cuda_device, geom):
    N, H1, W1 = geom
    Wp, RP = W1 + 2, (H1 + 2) * (W1 + 2)
    x, w, b, dy = _case(geom, 3)
    xp, dyp = _padded(x), _padded(dy)
    dwt = torch.zeros(384, 64, dtype=torch.float32, device=DEV)
    TO.call("conv3x3_dw_c32_c64", xp, xp.shape[0], dyp, dyp.shape[0], dwt, N * RP, Wp + 1, Wp)
    torch.cuda.synchronize()
    got = dwt.cpu().view(3, 4, 32, 64)
    assert got[:, 3].abs().max().item() == 0.0, "the unused fourth position slot must stay zero"
    got = got[:, :3].permute(3, 2, 0, 1)                            # [o, c, kh, kw]
    wr = w.clone().requires_grad_(True)
    F.conv2d(x, wr, b, padding=1).backward(dy)
    _assert_close(got, wr.grad, 2e-5, "weight gradient")          # exact bf16 products, fp32 accumulation in another order
    # accumulates: a second call doubles the sums
    TO.call("conv3x3_dw_c32_c64", xp, xp.shape[0], dyp, dyp.shape[0], dwt, N * RP, Wp + 1, Wp)
    torch.cuda.synchronize()
    _assert_close(dwt.cpu().view(3, 4, 32, 64)[:, :3].permute(3, 2, 0, 1), 2 * wr.grad, 2e-5, "weight gradient, accumulated")


def test_conv3x3_rejects_rows_wider_than_a_tile(cuda_device):
    W1 = 62                                                         # 128 + 2 * 64 + 2 input rows per tile > 256
    Wp, RP = W1 + 2, 3 * (W1 + 2)
    xp = torch.zeros(RP + 2 * Wp + 8, 32, dtype=torch.bfloat16, device=DEV)
    wseg = torch.zeros(64, 384, dtype=torch.bfloat16, device=DEV)
    y = torch.zeros((RP + 2 * Wp + 8) * 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="conv3x3"):
        TO.call("conv3x3_c32_c64", xp, xp.shape[0], wseg, torch.zeros(64, device=DEV), y, RP, Wp + 1, Wp)
