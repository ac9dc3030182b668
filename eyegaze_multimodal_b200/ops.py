"""Host side of the hot path: autograd Functions that launch the sm_100a kernels through the C ABI.

Every Function here enqueues kernels of libeyegaze_b200.so on the current CUDA stream through the ``torch.library`` ops
``torch.ops.eyegaze_b200.*`` (torch_ops.py: one op per entry point of include/eyegaze_b200.h, tensors in place of device
pointers); PyTorch only provides device memory, streams and the autograd graph between the fused ops.  Activations are stored in the *compute dtype* (fp32 for the
parity mode, bf16 for the throughput mode); parameters stay fp32 ``nn.Parameter``s and are re-cast once
per parameter version.  Nothing in this file has a CPU or eager fallback.
"""
import ctypes as C
import math
import threading

import os

import torch
from torch.utils.weak import WeakIdKeyDictionary

from . import _lib as L
from . import torch_ops as TO
from .precision import get_precision

F32, BF16 = L.F32, L.BF16
_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}


# ------------------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------------------
def _stream():
    return torch.cuda.current_stream().cuda_stream


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("unsupported dtype %s (fp32 / bf16 only)" % t.dtype)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("eyegaze_multimodal_b200 runs on CUDA tensors only (got a %s tensor); "
                               "there is no CPU path" % t.device)


def _p(t):
    return t          # (the registered ops take optional tensors where the C ABI takes nullable pointers)


def _rows(t: torch.Tensor):
    """(tensor, total_rows, rows_per_group, row_stride, group_stride) of a [.., K] view whose last dim is contiguous."""
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        t = t.contiguous()
    if t.dim() > 3:
        t = t.reshape(-1, t.shape[-2], t.shape[-1]) if t.is_contiguous() else t.contiguous().view(-1, t.shape[-2], t.shape[-1])
    if t.dim() == 2:
        return t, t.shape[0], 0, t.stride(0), 0
    G, R = t.shape[0], t.shape[1]
    if G == 1 or t.stride(0) == R * t.stride(1):
        return t, G * R, 0, t.stride(1), 0
    return t, G * R, R, t.stride(1), t.stride(0)


def _operand(t, major):
    t, rows, rpg, rs, gs = _rows(t)
    return TO.Operand(t, major, rpg, rs, gs, 0, 0), t


def _matrix(t):
    if t is None:
        return TO.Matrix(None, 0, 0, 0, 0), None
    t, rows, rpg, rs, gs = _rows(t)
    return TO.Matrix(t, _code(t), rpg, rs, gs), t


def _dense_matrix(t, code, ld):
    """Dense [rows, ld] matrix starting at the first element of tensor (view) ``t``."""
    return TO.Matrix(t, code, 0, ld, 0)


def _view_at(t, elem_offset=0):
    """The tensor form of ``t.data_ptr() + elem_offset * itemsize`` (strided sources included)."""
    if t.is_contiguous():
        return TO.at(t, elem_offset)
    return torch.as_strided(t, (1,), (1,), t.storage_offset() + elem_offset)


def zeros(shape, dtype, device):
    t = torch.empty(shape, dtype=dtype, device=device)
    TO.call("zero", t, t.numel() * t.element_size())
    return t


# Small fp32 accumulators (bias / LayerNorm parameter gradients: a few KB each, ~130 per training step) are carved out
# of 1 MB blocks that are zeroed with ONE memset each, instead of one memset launch per accumulator.  A block belongs to
# the stream it was zeroed on (the EEG branch runs on a side stream); views keep their block alive.
_ARENA_FLOATS = 1 << 18
_arenas = {}
_arena_lock = threading.Lock()


def small_zeros(shape, device):
    n = 1
    for d in shape:
        n *= int(d)
    if n > _ARENA_FLOATS // 8:
        return zeros(shape, torch.float32, device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    need = (n + 63) // 64 * 64                       # 256-byte granules keep every slice 16-byte aligned
    with _arena_lock:                                # backward runs on autograd's threads
        blk = _arenas.get(key)
        if blk is None or blk[1] + need > _ARENA_FLOATS:
            if len(_arenas) > 64:
                _arenas.clear()
            blk = [torch.zeros(_ARENA_FLOATS, dtype=torch.float32, device=device), 0]
            _arenas[key] = blk
        off = blk[1]
        blk[1] += need
    return blk[0][off:off + n].view(shape)


def reset_arenas() -> None:
    """Forget the partially used accumulator blocks.  Called when a CUDA-graph capture begins (a block zeroed BEFORE the
    capture would not be re-zeroed by replays) and when it ends (blocks allocated during capture belong to the graph)."""
    with _arena_lock:
        _arenas.clear()


_seed_epoch_word = {}


def enable_seed_epoch() -> int:
    """Creates the device word all dropout seeds are mixed with (include/eyegaze_b200.h: egb_seed_epoch_enable) and
    returns its address.  Needed once before a training step is captured into a CUDA graph."""
    dev = torch.cuda.current_device()
    if dev not in _seed_epoch_word:
        out = L.vp()
        L.call("egb_seed_epoch_enable", C.byref(out))       # host-side function (allocates the word): not an op
        _seed_epoch_word[dev] = out.value
    return _seed_epoch_word[dev]


def advance_seed_epoch() -> None:
    """One 1-thread launch on the current stream: the next kernels draw fresh dropout masks (capturable)."""
    enable_seed_epoch()
    TO.call("seed_epoch_advance")


def gemm(M, N, K, in_code, a: TO.Operand, b: TO.Operand, c: TO.Matrix, *, bias=None, c_pre=None, residual=None, aux=None,
         alpha=1.0, act=0, act_bwd=0, aux_scale=1.0, dropout_p=0.0, seed=0, accumulate=0, split_k=0, c_colsum=None):
    if in_code == F32 and not accumulate and split_k == 0 and get_precision() == "fp32":
        # parity mode: the K loop of a small fp32 GEMM stays in one CTA, in order -- the library would otherwise split it over
        # a thread-block cluster, an equally valid fp32 sum that can flip a ReLU at a pre-activation within 1e-6 of zero
        # (one element of one gradient of the default EEG model: 1 % of that tensor's max, twice the parity tolerance)
        split_k = 1
    empty = TO.Matrix(None, 0, 0, 0, 0)
    d = TO.GemmDesc(M, N, K, in_code, a, b, c, c_pre or empty, residual or empty, aux or empty, bias, alpha, act,
                   act_bwd, aux_scale, float(dropout_p), int(seed), accumulate, split_k, c_colsum)
    TO.call("gemm", d)


def _cast_raw(x: torch.Tensor, code: int) -> torch.Tensor:
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=_TORCH_DT[code], device=x.device)
    if code == F32:
        TO.call("cast_to_f32", x, _code(x), out, x.numel())
    else:
        TO.call("cast_from_f32", x, out, code, x.numel())
    return out


class CastFn(torch.autograd.Function):
    """dtype conversion that stays on the autograd graph (the gradient is converted back)."""

    @staticmethod
    def forward(ctx, x, code):
        ctx.src_code = _code(x)
        return _cast_raw(x, code)

    @staticmethod
    def backward(ctx, g):
        return (g if _code(g) == ctx.src_code else _cast_raw(g, ctx.src_code)), None


def cast(x: torch.Tensor, code: int) -> torch.Tensor:
    """dtype conversion through the library's cast kernels (fp32 <-> bf16); differentiable."""
    if _code(x) == code:
        return x
    if torch.is_grad_enabled() and x.requires_grad:
        return CastFn.apply(x, code)
    return _cast_raw(x, code)


def copy_strided4(src, dst, sizes, src_strides, dst_strides, src_offset=0, dst_offset=0):
    sz, ss, ds = [int(v) for v in sizes], [int(v) for v in src_strides], [int(v) for v in dst_strides]
    TO.call("copy_strided4", _view_at(src, src_offset), _code(src), _view_at(dst, dst_offset), _code(dst), sz, ss, ds)


def colsum(x: torch.Tensor, N: int) -> torch.Tensor:
    m, xt = _matrix(x)
    rows = _rows(xt)[1]
    out = small_zeros((N,), x.device)
    TO.call("colsum", m, rows, N, out, 0)
    return out


# dropout seeds: one fresh 64-bit value per op instance, derived from torch's seed so runs are reproducible
_seed_lock = threading.Lock()
_seed_state = {"base": None, "ctr": 0}


def next_seed() -> int:
    with _seed_lock:
        base = torch.initial_seed()
        if _seed_state["base"] != base:
            _seed_state["base"], _seed_state["ctr"] = base, 0
        _seed_state["ctr"] += 1
        x = (base * 0x9E3779B97F4A7C15 + _seed_state["ctr"] * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
        return x & 0x7FFFFFFFFFFFFFFF       # 63 bits: travels through int64 op arguments


# ------------------------------------------------------------------------------------------------------
# per-version cache of re-laid-out / re-cast parameter copies
# ------------------------------------------------------------------------------------------------------
class _WeightCache:
    """Derived copies live in a weak-id dictionary on the first source parameter, so an entry dies with its
    parameter and can never be served for a different tensor that happens to reuse the same address."""

    def __init__(self):
        self._d = WeakIdKeyDictionary()

    def get(self, params, code, tag, builder):
        p0 = params[0]
        slot = self._d.get(p0)
        if slot is None:
            slot = {}
            self._d[p0] = slot
        key = (code, tag, tuple(id(p) for p in params[1:]))
        ver = tuple(p._version for p in params) + tuple(p.data_ptr() for p in params) + (_param_epoch[0],)
        hit = slot.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = builder()
        slot[key] = (ver, val)
        return val


wcache = _WeightCache()
_param_epoch = [0]


def bump_param_epoch() -> None:
    """Invalidate every derived parameter copy (bf16 casts, packed QKV): called by code that updates parameters
    without going through autograd's version counters (optim.FusedClipAdamW)."""
    _param_epoch[0] += 1


# bf16 copies of fp32 parameters live in PERSISTENT buffers (one per parameter), so that all of them can be re-derived by
# one multi-tensor launch (refresh_plain_copies) after an optimizer update instead of one cast launch per parameter.
_plain_bufs = WeakIdKeyDictionary()          # parameter -> bf16 buffer [N, K]
_plain_tables = {"sig": None, "cast": None, "chunks": None, "n_chunks": 0, "items": []}
_PLAIN_CHUNK = 16384


def _plain_build(w: torch.Tensor, w2: torch.Tensor, code: int) -> torch.Tensor:
    if code != BF16 or w2.dtype != torch.float32 or not w2.is_contiguous():
        return cast(w2, code)
    buf = _plain_bufs.get(w)
    if buf is None or buf.shape != w2.shape or buf.device != w2.device:
        buf = torch.empty(w2.shape, dtype=torch.bfloat16, device=w2.device)
        _plain_bufs[w] = buf
    TO.call("cast_from_f32", w2, buf, code, w2.numel())
    return buf


def weight_plain(w: torch.Tensor, code: int) -> torch.Tensor:
    """[N, K...] parameter flattened to [N, K] in the compute dtype (no copy in fp32 mode)."""
    w2 = w.detach().reshape(w.shape[0], -1)
    if code == F32:
        return w2
    return wcache.get((w,), code, "plain", lambda: _plain_build(w, w2, code))


def refresh_plain_copies(device=None) -> int:
    """Re-derive EVERY registered bf16 weight copy from its fp32 master with one launch and mark the cache entries valid
    for the current parameter versions / epoch.  The (static) pointer tables are built on the first call and whenever the
    set of copies changes -- that upload cannot happen during stream capture, so a captured step calls this once eagerly
    after its warm-up.  Returns the number of tensors refreshed."""
    import numpy as np
    if not torch.cuda.is_available():
        return 0
    dev_want = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev_want.index is None:
        dev_want = torch.device("cuda", torch.cuda.current_device())
    items = []
    for w, buf in list(_plain_bufs.items()):
        if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.device == dev_want:   # one GPU per process
            items.append((w, buf))
    if not items:
        return 0
    sig = tuple((w.data_ptr(), buf.data_ptr(), w.numel()) for w, buf in items)
    T = _plain_tables
    if T["sig"] != sig:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("refresh_plain_copies: the set of weight copies changed during stream capture")
        dev = items[0][0].device
        rec = np.zeros(len(items), dtype=[("src", "<u8"), ("dst", "<u8")])
        rec["src"] = [w.data_ptr() for w, _ in items]
        rec["dst"] = [b.data_ptr() for _, b in items]
        chunks = []
        for i, (w, _) in enumerate(items):
            for o in range(0, w.numel(), _PLAIN_CHUNK):
                chunks.append((i, min(_PLAIN_CHUNK, w.numel() - o), o))
        ch = np.zeros(len(chunks), dtype=[("tensor", "<i4"), ("n", "<i4"), ("offset", "<i8")])
        ch["tensor"], ch["n"], ch["offset"] = zip(*chunks)
        T["cast"] = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
        T["chunks"] = torch.from_numpy(ch.view(np.uint8).copy()).to(dev)
        T["n_chunks"] = len(chunks)
        T["sig"] = sig
    TO.call("multi_tensor_cast_bf16", T["cast"], T["chunks"], T["n_chunks"])
    for w, buf in items:
        slot = wcache._d.get(w)
        if slot is None:
            slot = {}
            wcache._d[w] = slot
        slot[(BF16, "plain", ())] = ((w._version, w.data_ptr(), _param_epoch[0]), buf)
    return len(items)


def weight_packed(ws, code):
    """Row-concatenation of several [N_i, K] parameters (QKV packing) in the compute dtype."""
    def build():
        n = sum(w.shape[0] for w in ws)
        out = torch.empty(n, ws[0].shape[1], dtype=_TORCH_DT[code], device=ws[0].device)
        r = 0
        for w in ws:
            copy_strided4(w.detach(), out, (1, 1, w.shape[0], w.shape[1]), (0, 0, w.shape[1], 1), (0, 0, w.shape[1], 1),
                          dst_offset=r * w.shape[1])
            r += w.shape[0]
        return out
    return wcache.get(tuple(ws), code, "packed", build)


def bias_packed(bs):
    def build():
        n = sum(b.shape[0] for b in bs)
        out = torch.empty(n, dtype=torch.float32, device=bs[0].device)
        r = 0
        for b in bs:
            copy_strided4(b.detach(), out, (1, 1, 1, b.shape[0]), (0, 0, 0, 1), (0, 0, 0, 1), dst_offset=r)
            r += b.shape[0]
        return out
    return wcache.get(tuple(bs), F32, "bias_packed", build)


# ------------------------------------------------------------------------------------------------------
# Linear (+bias +relu +dropout +residual)
# ------------------------------------------------------------------------------------------------------
def _linear_forward(x, w2, bias, residual, act, p, seed, out_code, c_pre=None):
    """x: [.., K] (compute dtype), w2: [N, K] same dtype.  Returns y [.., N]."""
    a, xt = _operand(x, 0)
    rows = _rows(xt)[1]
    N, K = w2.shape
    y = torch.empty(tuple(x.shape[:-1]) + (N,), dtype=_TORCH_DT[out_code], device=x.device)
    cm = _dense_matrix(y, out_code, N)
    rm, rt = _matrix(residual)
    pm = _dense_matrix(c_pre, _code(c_pre), N) if c_pre is not None else None
    gemm(rows, N, K, _code(xt), a, TO.Operand(w2, 0, 0, w2.stride(0), 0, 0, 0), cm, bias=bias, residual=rm if rt is not None else None,
         c_pre=pm, act=act, dropout_p=p, seed=seed)
    return y


def _grad_input(dpre, w2, x_shape, out_code, act_bwd=0, aux=None, aux_scale=1.0, dropout_p=0.0, seed=0, colsum_out=None):
    """dx[.., K] = dpre[.., N] . W[N, K]  (W consumed as an MN-major operand, no transposed copy).
    ``colsum_out`` ([K] fp32, zeroed by the caller) receives the column sums of dx -- the bias gradient of the layer
    that produced x -- from the GEMM epilogue."""
    a, dt = _operand(dpre, 0)
    rows = _rows(dt)[1]
    N, K = w2.shape
    dx = torch.empty(tuple(x_shape), dtype=_TORCH_DT[out_code], device=dpre.device)
    am = _dense_matrix(aux, _code(aux), K) if aux is not None else None
    gemm(rows, K, N, _code(dt), a, TO.Operand(w2, 1, 0, w2.stride(0), 0, 0, 0),
         _dense_matrix(dx, out_code, K), act_bwd=act_bwd, aux=am, aux_scale=aux_scale, dropout_p=dropout_p,
         seed=seed, c_colsum=colsum_out)
    return dx


def _grad_weight(dpre, x, N, K):
    """dW[N, K] (fp32) = dpre^T . x, split-K accumulation; both operands MN-major views (no transposes)."""
    a, dt = _operand(dpre, 1)
    b, xt = _operand(x, 1)
    rows = _rows(dt)[1]
    dw = torch.empty(N, K, dtype=torch.float32, device=dpre.device)
    gemm(N, K, rows, _code(dt), a, b, _dense_matrix(dw, F32, K), accumulate=2)
    return dw


def _act_bwd(dy, aux, mode, scale):
    """dense dpre = dy * act'(aux); dy may be a strided view."""
    dm, dt = _matrix(dy)
    rows = _rows(dt)[1]
    N = dt.shape[-1]
    out = torch.empty(tuple(dy.shape), dtype=dt.dtype, device=dy.device)
    om = _dense_matrix(out, _code(out), N)
    TO.call("act_bwd", dm, aux, om, rows, N, mode, float(scale))
    return out


def _dropout_bwd(dy, p, seed):
    dy = dy.contiguous()
    out = torch.empty_like(dy)
    TO.call("dropout_bwd", dy, out, _code(dy), dy.numel(), float(p), int(seed))
    return out


def _dropout_bwd_colsum(dy, p, seed, N):
    """-> (dy * mask / (1-p), its column sums [N] fp32) in ONE pass (the bias gradient of the layer whose epilogue applied
    the mask); falls back to two passes for row widths the fused kernel does not take."""
    dy = dy.contiguous()
    if N % 8 != 0 or N > 1024 or dy.shape[-1] != N:
        out = _dropout_bwd(dy, p, seed)
        return out, colsum(out, N)
    out = torch.empty_like(dy)
    cs = small_zeros((N,), dy.device)
    TO.call("dropout_bwd_colsum", dy, out, _code(dy), dy.numel() // N, N, float(p), int(seed),
           cs)
    stats["colsum_fused"] += 1
    return out, cs


# Bias gradients without a pass of their own: a kernel that PRODUCES a gradient tensor (LayerNorm backward, attention
# backward) also accumulates its column sums and hands them along as an attribute of that tensor object; the Linear whose
# output gradient it is picks them up.  The hint dies with the tensor object, so a gradient that autograd re-creates
# (accumulation of several consumers, a dtype cast) simply has none and the column-sum pass runs as before.
stats = {"colsum_fused": 0, "colsum_pass": 0}


def _attach_colsum(t, cs):
    t._egb_colsum = (cs, t._version)       # the version pins the hint to the tensor's CONTENT (autograd may add in place)
    return t


def _bias_grad(dpre, N):
    hint = getattr(dpre, "_egb_colsum", None)
    if hint is not None and hint[1] == dpre._version and hint[0].numel() == N and dpre.shape[-1] == N:
        stats["colsum_fused"] += 1
        return hint[0]
    stats["colsum_pass"] += 1
    return colsum(dpre, N)


class LinearFn(torch.autograd.Function):
    """y = [residual +] dropout(act(x W^T + b)).  `weights`/`biases`: 1 parameter, or several packed along N."""

    @staticmethod
    def forward(ctx, x, residual, n_w, act, p, out_f32, *wb):
        weights, biases = wb[:n_w], wb[n_w:]
        code = _code(x)
        _require_cuda(x, *weights)
        if act != 0 and residual is not None:
            raise ValueError("linear: an activation combined with a residual is not a reference op chain")
        w2 = weight_plain(weights[0], code) if n_w == 1 else weight_packed(weights, code)
        has_bias = len(biases) > 0 and biases[0] is not None
        bias = None
        if has_bias:
            bias = biases[0].detach() if n_w == 1 else bias_packed([b.detach() for b in biases])
        seed = next_seed() if p > 0 else 0
        out_code = F32 if out_f32 else code
        y = _linear_forward(x, w2, bias, residual, act, p, seed, out_code)
        ctx.save_for_backward(x, y if act == L.ACT_RELU else None, *weights)
        ctx.meta = (n_w, act, p, seed, has_bias, residual is not None, code)
        return y

    @staticmethod
    def backward(ctx, dy):
        n_w, act, p, seed, has_bias, has_res, code = ctx.meta
        saved = ctx.saved_tensors
        x, y, weights = saved[0], saved[1], saved[2:]
        w2 = weight_plain(weights[0], code) if n_w == 1 else weight_packed(weights, code)
        N, K = w2.shape
        if _code(dy) != code:
            dy = cast(dy, code)
        want_b = has_bias and any(ctx.needs_input_grad[6 + n_w:])
        db = None
        if act == L.ACT_RELU:
            dpre = _act_bwd(dy, y, 1, 1.0 / (1.0 - p) if p > 0 else 1.0)
        elif p > 0:
            if want_b:
                dpre, db = _dropout_bwd_colsum(dy, p, seed, N)
            else:
                dpre = _dropout_bwd(dy, p, seed)
        else:
            dpre = dy
        dx = _grad_input(dpre, w2, x.shape, code) if ctx.needs_input_grad[0] else None
        dres = dy if has_res else None
        need_w = any(ctx.needs_input_grad[6 + i] for i in range(n_w))
        dws = [None] * n_w
        dbs = [None] * (len(ctx.needs_input_grad) - 6 - n_w)
        if need_w:
            dw = _grad_weight(dpre, x, N, K)
            r = 0
            for i, w in enumerate(weights):
                dws[i] = dw[r:r + w.shape[0]].view(w.shape)
                r += w.shape[0]
        if want_b:
            if db is None:
                db = _bias_grad(dpre, N)
            r = 0
            for i, w in enumerate(weights):
                dbs[i] = db[r:r + w.shape[0]]
                r += w.shape[0]
        return (dx, dres, None, None, None, None) + tuple(dws) + tuple(dbs)


def linear(x, weight, bias=None, residual=None, act=0, p=0.0, out_f32=False):
    return LinearFn.apply(x, residual, 1, act, float(p), out_f32, weight, bias)


def linear_packed(x, weights, biases, out_f32=False):
    return LinearFn.apply(x, None, len(weights), 0, 0.0, out_f32, *weights, *biases)


# ------------------------------------------------------------------------------------------------------
# two-layer MLP:  y = [residual +] drop_out(act(drop_mid(x W1^T + b1)) W2^T + b2)     (fused backward)
# ------------------------------------------------------------------------------------------------------
class Mlp2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, w1, b1, w2, b2, act, p_mid, p_out, out_f32):
        code = _code(x)
        _require_cuda(x, w1, w2)
        w1c, w2c = weight_plain(w1, code), weight_plain(w2, code)
        s_mid = next_seed() if p_mid > 0 else 0
        s_out = next_seed() if p_out > 0 else 0
        pre = None
        if act == L.ACT_GELU:
            pre = torch.empty(tuple(x.shape[:-1]) + (w1.shape[0],), dtype=x.dtype, device=x.device)
        # GELU: the epilogue saves gelu'(pre) (not pre), so the backward epilogue is a plain multiply
        h = _linear_forward(x, w1c, b1.detach(), None, L.ACT_GELU_DGRAD if act == L.ACT_GELU else act, p_mid, s_mid, code,
                            c_pre=pre)
        y = _linear_forward(h, w2c, b2.detach(), residual, 0, p_out, s_out, F32 if out_f32 else code)
        ctx.save_for_backward(x, h, pre, w1, w2)
        ctx.meta = (act, p_mid, p_out, s_mid, s_out, residual is not None, code)
        return y

    @staticmethod
    def backward(ctx, dy):
        act, p_mid, p_out, s_mid, s_out, has_res, code = ctx.meta
        x, h, pre, w1, w2 = ctx.saved_tensors
        w1c, w2c = weight_plain(w1, code), weight_plain(w2, code)
        if _code(dy) != code:
            dy = cast(dy, code)
        need = ctx.needs_input_grad
        db2 = None
        if p_out > 0 and need[5]:
            dyd, db2 = _dropout_bwd_colsum(dy, p_out, s_out, w2.shape[0])
        else:
            dyd = _dropout_bwd(dy, p_out, s_out) if p_out > 0 else dy
            if need[5]:
                db2 = _bias_grad(dyd, w2.shape[0])
        dw2 = _grad_weight(dyd, h, w2.shape[0], w2.shape[1]) if need[4] else None
        # dpre1 = (dyd . W2) * act'(.)   -- activation derivative, mid-dropout mask AND the first layer's bias gradient
        # (column sums of dpre1) fused in the GEMM epilogue
        db1 = small_zeros((w1.shape[0],), dy.device) if need[3] else None
        if act == L.ACT_RELU:
            dpre1 = _grad_input(dyd, w2c, h.shape, code, act_bwd=L.ACTBWD_RELU_MASK, aux=h,
                                aux_scale=1.0 / (1.0 - p_mid) if p_mid > 0 else 1.0, colsum_out=db1)
        elif act == L.ACT_GELU:
            dpre1 = _grad_input(dyd, w2c, h.shape, code, act_bwd=L.ACTBWD_MUL, aux=pre, dropout_p=p_mid, seed=s_mid,
                                colsum_out=db1)
        else:
            dpre1 = _grad_input(dyd, w2c, h.shape, code, dropout_p=p_mid, seed=s_mid, colsum_out=db1)
        dw1 = _grad_weight(dpre1, x, w1.shape[0], w1.shape[1]) if need[2] else None
        dx = _grad_input(dpre1, w1c, x.shape, code) if need[0] else None
        return dx, (dy if has_res else None), dw1, db1, dw2, db2, None, None, None, None


def mlp2(x, w1, b1, w2, b2, act, p_mid=0.0, p_out=0.0, residual=None, out_f32=False):
    return Mlp2Fn.apply(x, residual, w1, b1, w2, b2, act, float(p_mid), float(p_out), out_f32)


# ------------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _require_cuda(x, gamma)
        x = x.contiguous()
        D = x.shape[-1]
        M = x.numel() // D
        y = torch.empty_like(x)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
        TO.call("layernorm_fwd", x, gamma, beta, y, mean,
               rstd, _code(x), M, D, float(eps))
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        D = x.shape[-1]
        M = x.numel() // D
        dy = dy.contiguous()
        if _code(dy) != _code(x):
            dy = cast(dy, _code(x))
        dx = torch.empty_like(x)
        dgb = small_zeros((3, D), x.device)        # dgamma | dbeta | column sums of dx (bias gradient of the producer)
        TO.call("layernorm_bwd_ex", dy, x, gamma, mean, rstd,
               dx, dgb[0], dgb[1], None, dgb[2], _code(x), M, D)
        return _attach_colsum(dx, dgb[2]), dgb[0], dgb[1], None


def layernorm(x, gamma, beta, eps):
    return LayerNormFn.apply(x, gamma, beta, float(eps))


class LayerNormResidualFn(torch.autograd.Function):
    """(LayerNorm(x), x) for a pre-norm residual block: the second output is x itself and is meant to be used as
    the block's residual input.  Backward receives the gradients of both uses of x at once and adds them inside the
    LayerNorm backward kernel, instead of autograd summing two [tokens, D] tensors in a separate pass."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _require_cuda(x, gamma)
        x = x.contiguous()
        D = x.shape[-1]
        M = x.numel() // D
        y = torch.empty_like(x)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
        TO.call("layernorm_fwd", x, gamma, beta, y, mean,
               rstd, _code(x), M, D, float(eps))
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x, gamma, mean, rstd = ctx.saved_tensors
        D = x.shape[-1]
        M = x.numel() // D
        code = _code(x)
        dgb = small_zeros((3, D), x.device)
        if dy is None:                      # only the residual path was used
            return dres, dgb[0], dgb[1], None
        dy = dy.contiguous()
        if _code(dy) != code:
            dy = cast(dy, code)
        if dres is not None:
            dres = dres.contiguous()
            if _code(dres) != code:
                dres = cast(dres, code)
        dx = torch.empty_like(x)
        TO.call("layernorm_bwd_ex", dy, x, gamma, mean, rstd,
               dx, dgb[0], dgb[1], _p(dres), dgb[2], code, M, D)
        return _attach_colsum(dx, dgb[2]), dgb[0], dgb[1], None


def layernorm_residual(x, gamma, beta, eps):
    """-> (LayerNorm(x), x as the residual operand); see LayerNormResidualFn."""
    if not (torch.is_grad_enabled() and (x.requires_grad or gamma.requires_grad)):
        return LayerNormFn.apply(x, gamma, beta, float(eps)), x
    return LayerNormResidualFn.apply(x, gamma, beta, float(eps))


# ------------------------------------------------------------------------------------------------------
# fused attention
# ------------------------------------------------------------------------------------------------------
import os as _os
_ATT_COLSUM = _os.environ.get("EGB_ATT_COLSUM", "1") != "0"     # 0: bias gradient of qkv by a separate column-sum pass


def _att_strides(t):
    """(batch_stride, row_stride) of a [S, L, H*dk] view with contiguous last dim."""
    return t.stride(0), t.stride(1)


class AttentionFn(torch.autograd.Function):
    """ctx = softmax(q k^T * scale) [dropout] v over heads.  q: [S, Lq, D], k/v: [S, Lk, D] (views allowed).

    With ``packed=True`` the single input is the packed projection [S, L, 3D] (q | k | v) and a single packed
    gradient is returned.  kv_shift pairs query batch s with key/value batch (s + kv_shift) % S.
    """

    @staticmethod
    def forward(ctx, q, k, v, heads, kv_shift, p, scale, packed, want_probs):
        _require_cuda(q)
        if packed:
            qkv = q.contiguous()
            D = qkv.shape[-1] // 3
            q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        else:
            q = q if q.stride(-1) == 1 else q.contiguous()
            k = k if k.stride(-1) == 1 else k.contiguous()
            v = v if v.stride(-1) == 1 else v.contiguous()
            D = q.shape[-1]
        S, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
        dk = D // heads
        code = _code(q)
        o = torch.empty(S, Lq, D, dtype=q.dtype, device=q.device)
        lse = torch.empty(S, heads, Lq, dtype=torch.float32, device=q.device)
        probs = torch.empty(S, heads, Lq, Lk, dtype=torch.float32, device=q.device) if want_probs else None
        seed = next_seed() if p > 0 else 0
        d = TO.AttentionDesc()
        d.q, d.k, d.v, d.o = q, k, v, o
        d.q_bs, d.q_rs = _att_strides(q)
        d.k_bs, d.k_rs = _att_strides(k)
        d.v_bs, d.v_rs = _att_strides(v)
        d.o_bs, d.o_rs = _att_strides(o)
        d.lse, d.probs = lse, probs
        d.dtype, d.S, d.H, d.Lq, d.Lk, d.head_dim, d.kv_shift = code, S, heads, Lq, Lk, dk, kv_shift
        d.scale, d.dropout_p, d.seed = float(scale), float(p), seed
        TO.call("attention_fwd", d)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (heads, kv_shift, p, scale, seed, packed)
        if want_probs:
            ctx.mark_non_differentiable(probs)
            return o, probs
        return o

    @staticmethod
    def backward(ctx, do, *unused):
        heads, kv_shift, p, scale, seed, packed = ctx.meta
        q, k, v, o, lse = ctx.saved_tensors
        S, Lq, D = o.shape
        Lk = k.shape[1]
        code = _code(q)
        do = do if do.stride(-1) == 1 else do.contiguous()
        if _code(do) != code:
            do = cast(do, code)
        if packed:
            dqkv = torch.empty(S, Lq, 3 * D, dtype=q.dtype, device=q.device)
            dq, dk_, dv = dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:]
        else:
            dq = torch.empty(S, Lq, D, dtype=q.dtype, device=q.device)
            dk_ = torch.empty(S, Lk, D, dtype=q.dtype, device=q.device)
            dv = torch.empty(S, Lk, D, dtype=q.dtype, device=q.device)
        delta = torch.empty(S, heads, Lq, dtype=torch.float32, device=q.device)
        d = TO.AttentionDesc()
        d.q, d.k, d.v, d.o = q, k, v, o
        d.q_bs, d.q_rs = _att_strides(q)
        d.k_bs, d.k_rs = _att_strides(k)
        d.v_bs, d.v_rs = _att_strides(v)
        d.o_bs, d.o_rs = _att_strides(o)
        d.d_o = do
        d.do_bs, d.do_rs = _att_strides(do)
        d.dq, d.dk, d.dv = dq, dk_, dv
        d.dq_bs, d.dq_rs = _att_strides(dq)
        d.dk_bs, d.dk_rs = _att_strides(dk_)
        d.dv_bs, d.dv_rs = _att_strides(dv)
        d.lse, d.delta, d.probs = lse, delta, None
        d.dtype, d.S, d.H, d.Lq, d.Lk, d.head_dim, d.kv_shift = code, S, heads, Lq, Lk, D // heads, kv_shift
        d.scale, d.dropout_p, d.seed = float(scale), float(p), seed
        cs = None
        if packed and _ATT_COLSUM:
            # column sums of dq | dk | dv = the bias gradient of the packed qkv projection, taken as the tiles leave the kernel
            cs = small_zeros((3 * D,), q.device)
            d.dq_colsum, d.dk_colsum, d.dv_colsum = cs[:D], cs[D:2 * D], cs[2 * D:]
        TO.call("attention_bwd", d)
        if packed:
            return (_attach_colsum(dqkv, cs) if cs is not None else dqkv), None, None, None, None, None, None, None, None
        return dq, dk_, dv, None, None, None, None, None, None


def attention_packed(qkv, heads, kv_shift=0, p=0.0, scale=None, want_probs=False):
    D = qkv.shape[-1] // 3
    scale = scale if scale is not None else 1.0 / math.sqrt(D // heads)
    return AttentionFn.apply(qkv, None, None, heads, kv_shift, float(p), scale, True, want_probs)


def attention(q, k, v, heads, p=0.0, scale=None, want_probs=False):
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1] // heads)
    return AttentionFn.apply(q, k, v, heads, 0, float(p), scale, False, want_probs)


# ------------------------------------------------------------------------------------------------------
# temporal-conv frontend: [Conv1d(k, stride, pad=k//2) -> ReLU -> Dropout] x n as implicit GEMMs
# ------------------------------------------------------------------------------------------------------
def _conv_weight(w, code):
    """(O, C, K) parameter -> [O, K*C] (tap-major, channel-minor: matches the channels-last operand view)."""
    O, Cc, K = w.shape

    def build():
        out = torch.empty(O, K * Cc, dtype=_TORCH_DT[code], device=w.device)
        copy_strided4(w.detach(), out, (1, O, K, Cc), (0, Cc * K, 1, K), (0, K * Cc, Cc, 1))
        return out
    return wcache.get((w,), code, "conv1d", build)


def _conv_weight_phase(w, code, stride, J):
    """Transposed-conv operand: Bp[p][c, (jj, o)] = W[o, c, p + stride*(J-1-jj)] (0 where the tap does not exist)."""
    O, Cc, K = w.shape

    def build():
        out = zeros((stride, Cc, J * O), _TORCH_DT[code], w.device)
        for ph in range(stride):
            for jj in range(J):
                tap = ph + stride * (J - 1 - jj)
                if tap < K:
                    copy_strided4(w.detach(), out, (1, 1, Cc, O), (0, 0, K, Cc * K), (0, 0, J * O, 1), src_offset=tap,
                                  dst_offset=ph * Cc * J * O + jj * O)
        return out
    return wcache.get((w,), code, "conv1d_phase", build)


_debug = None   # set to a dict by debugging scripts to capture backward intermediates


def _round_up(x, m):
    return (x + m - 1) // m * m


class TemporalConvFn(torch.autograd.Function):
    """eeg1, eeg2 (B,C,T) fp32 -> h [2B, T~, d] (players stacked along the batch).  dual_eeg_transformer.py:163-175."""

    @staticmethod
    def forward(ctx, eeg1, eeg2, code, stride, p, *wb):
        n = len(wb) // 2
        ws, bs = wb[:n], wb[n:]
        _require_cuda(eeg1, eeg2, *ws)
        eeg1 = eeg1.contiguous().float()
        eeg2 = eeg2.contiguous().float()
        B, Cin, T = eeg1.shape
        S = 2 * B
        dev = eeg1.device
        tdt = _TORCH_DT[code]
        ksz = ws[0].shape[2]
        pad = ksz // 2
        if n > 1 and any(w.requires_grad for w in ws) and (stride < 2 or pad % stride != 0 or pad < stride - 1):
            # the input-gradient of layers >= 1 runs as `stride` phase GEMMs whose slab geometry needs these
            raise NotImplementedError("temporal conv frontend: training needs stride >= 2 and (kernel_size // 2) %% stride "
                                      "== 0 (got kernel_size=%d, stride=%d); the reference configs use 25 / 4"
                                      % (ksz, stride))
        # channels-last, zero-padded input; group stride kept a multiple of 8 elements for TMA
        Tp = T + 2 * pad
        while (Tp * Cin) % 8:
            Tp += 1
        xp = torch.empty(S, Tp, Cin, dtype=tdt, device=dev)
        TO.call("eeg_pack", eeg1, eeg2, xp, code, B, Cin, T, pad, Tp)
        bufs, geo, seeds = [xp], [(T, Tp, Cin)], []
        cur, t_in, tp_in, c_in = xp, T, Tp, Cin
        for i in range(n):
            O = ws[i].shape[0]
            t_out = (t_in + 2 * pad - ksz) // stride + 1
            last = i == n - 1
            tp_out = t_out if last else t_out + 2 * pad
            out = torch.empty(S, tp_out, O, dtype=tdt, device=dev) if last else zeros((S, tp_out, O), tdt, dev)
            wr = _conv_weight(ws[i], code)
            seed = next_seed() if p > 0 else 0
            seeds.append(seed)
            a = TO.Operand(cur, 0, t_out, stride * c_in, tp_in * c_in, 0, 0)
            off = 0 if last else pad * O
            cm = TO.Matrix(TO.at(out, off), code, t_out, O, tp_out * O)
            gemm(S * t_out, O, ksz * c_in, code, a, TO.Operand(wr, 0, 0, wr.stride(0), 0, 0, 0), cm,
                 bias=bs[i].detach(), act=L.ACT_RELU, dropout_p=p, seed=seed)
            bufs.append(out)
            geo.append((t_out, tp_out, O))
            cur, t_in, tp_in, c_in = out, t_out, tp_out, O
        ctx.save_for_backward(*bufs, *ws)
        ctx.meta = (n, code, stride, p, ksz, pad, S, geo)
        return cur

    @staticmethod
    def backward(ctx, dh):
        n, code, stride, p, ksz, pad, S, geo = ctx.meta
        saved = ctx.saved_tensors
        bufs, ws = saved[:n + 1], saved[n + 1:]
        tdt = _TORCH_DT[code]
        dev = dh.device
        scale = 1.0 / (1.0 - p) if p > 0 else 1.0
        if _code(dh) != code:
            dh = cast(dh, code)
        J = (ksz + stride - 1) // stride
        if n > 1 and (pad % stride != 0 or stride < 2):
            raise NotImplementedError("temporal conv backward needs stride >= 2 and (kernel//2) %% stride == 0")
        dws, dbs = [None] * n, [None] * n
        # gradient of the last layer's pre-activation, written with J-1-s0 zero rows in front (slab view below)
        s0 = pad // stride
        front = max(J - 1 - s0, 0)
        t_out, _, O = geo[n]
        rp = front + t_out + J
        g = zeros((S, rp, O), tdt, dev)
        dm, dht = _matrix(dh)
        gm = TO.Matrix(TO.at(g, front * O), code, t_out, O, rp * O)
        TO.call("act_bwd", dm, bufs[n], gm, S * t_out, O, 1, float(scale))
        g_front, g_rp = front, rp
        if _debug is not None:
            _debug["g0"] = (g.clone(), front, rp)
        for i in range(n - 1, -1, -1):
            t_out, _, O = geo[i + 1]
            t_in, tp_in, c_in = geo[i]
            src = bufs[i]
            g_ptr = TO.at(g, g_front * O)
            # dW[O, K*C] = dpre^T . (overlapping-row view of the layer input); db = column sums
            if ctx.needs_input_grad[5 + i]:
                dw = torch.empty(O, ksz * c_in, dtype=torch.float32, device=dev)
                gemm(O, ksz * c_in, S * t_out, code, TO.Operand(g_ptr, 1, t_out, O, g_rp * O, 0, 0),
                     TO.Operand(src, 1, t_out, stride * c_in, tp_in * c_in, 0, 0),
                     _dense_matrix(dw, F32, ksz * c_in), accumulate=2)
                dws[i] = dw.view(O, ksz, c_in).permute(0, 2, 1)
            if ctx.needs_input_grad[5 + n + i]:
                db = torch.empty(O, dtype=torch.float32, device=dev)
                gmat = TO.Matrix(g_ptr, code, t_out, O, g_rp * O)
                TO.call("colsum", gmat, S * t_out, O, db, 1)
                dbs[i] = db
            if i == 0:
                break
            # d(input) via `stride` phase GEMMs (transposed convolution), fused with the previous ReLU(+dropout) mask.
            # t_in need not be a multiple of the stride: the last row of a phase may land up to stride-1 positions
            # past the sequence end, i.e. in the zero padding of the saved activation, where the ReLU mask (y != 0)
            # makes the stored gradient exactly 0 again.
            wp = _conv_weight_phase(ws[i], code, stride, J)
            g_prev = zeros((S, tp_in, c_in), tdt, dev)   # same padded geometry as bufs[i]
            rows = (t_in + stride - 1) // stride
            slab0 = g_front - (J - 1 - s0)
            for ph in range(stride):
                a = TO.Operand(TO.at(g, slab0 * O), 0, rows, O, g_rp * O, 0, 0)
                bop = TO.Operand(TO.at(wp, ph * c_in * J * O), 0, 0, J * O, 0, 0, 0)
                off = (pad + ph) * c_in
                cm = TO.Matrix(TO.at(g_prev, off), code, rows, stride * c_in, tp_in * c_in)
                am = TO.Matrix(TO.at(src, off), code, rows, stride * c_in, tp_in * c_in)
                gemm(S * rows, c_in, J * O, code, a, bop, cm, act_bwd=L.ACTBWD_RELU_MASK, aux=am, aux_scale=scale)
            g, g_front, g_rp = g_prev, pad, tp_in
            if _debug is not None:
                _debug["g%d" % (n - i)] = (g.clone(), pad, tp_in)
        return (None, None, None, None, None) + tuple(dws) + tuple(dbs)


def temporal_conv(eeg1, eeg2, weights, biases, code, stride, p):
    return TemporalConvFn.apply(eeg1, eeg2, code, stride, float(p), *weights, *biases)


# ------------------------------------------------------------------------------------------------------
# spectrogram CNN:  STFT-log -> conv(1->32)+relu+maxpool -> conv(32->64) [tensor-core implicit GEMM] -> relu+avgpool
# ------------------------------------------------------------------------------------------------------
def _spec_w2_seg(w, code):
    """(64, 32, 3, 3) -> [64, 3 segs x (4 kw x 32 c)] with a zero kw=3 slot (128-element K segments)."""
    def build():
        O, Cc = w.shape[0], w.shape[1]
        out = zeros((O, 3 * 4 * Cc), _TORCH_DT[code], w.device)
        for kh in range(3):
            copy_strided4(w.detach(), out, (1, O, 3, Cc), (0, Cc * 9, 1, 9), (0, 12 * Cc, Cc, 1), src_offset=kh * 3,
                          dst_offset=kh * 4 * Cc)
        return out
    return wcache.get((w,), code, "spec_w2_seg", build)


# kw slots per K segment of the conv-2 dX GEMM: 4 = a zero slot pads the segment to 256 (128-wide TMA stages); 3 = no padding,
# 192-wide segments, but those only fit 64-wide stages -- measured SLOWER (1.98 vs 1.57 ms): the per-instruction cost of TMA
# outweighs the 25 % of operand bytes saved
_SPEC_DX_SLOTS = int(os.environ.get("EGB_SPEC_DX_SLOTS", "4"))
_SPEC_DX_DIRECT = int(os.environ.get("EGB_SPEC_DX_DIRECT", "1"))    # 0: the generic GEMM over the overlapping-row view


def _spec_w2_flip(w, code):
    """dX operand: [32 c, 3 segs(a) x (S b x 64 o)] = W[o, c, 2-a, 2-b]; S = 3 (K segments of 192 = three positions x 64
    channels, no padding) or 4 (a zero b = 3 slot: 256-wide segments, 25 % of the GEMM spent on zeros)."""
    S = _SPEC_DX_SLOTS

    def build():
        O, Cc = w.shape[0], w.shape[1]
        out = zeros((Cc, 3 * S * O), _TORCH_DT[code], w.device)
        for a in range(3):
            for b in range(3):
                copy_strided4(w.detach(), out, (1, 1, Cc, O), (0, 0, 9, Cc * 9), (0, 0, 3 * S * O, 1),
                              src_offset=(2 - a) * 3 + (2 - b), dst_offset=a * S * O + b * O)
        return out
    return wcache.get((w,), code, "spec_w2_flip%d" % S, build)


def _spec_geometry(bins, frames):
    H1, W1 = bins // 2, frames // 2
    Wp, Hp = W1 + 2, H1 + 2
    return H1, W1, Wp, Hp * Wp, 2 * Wp + 8           # H1, W1, padded width, padded rows per image, slack rows


def _spec_front_fwd(eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins, want_amax=False):
    """STFT-log -> conv1+ReLU+maxpool -> conv2 (+bias).  Returns (img, p1, y2, meta); p1 / y2 are zero-bordered
    channels-last images, flat [(N*RP + slack) * C]; y2 is the spec_conv[3] output (pre-ReLU)."""
    eeg1 = eeg1.contiguous().float()
    eeg2 = eeg2.contiguous().float()
    B, Cc, T = eeg1.shape
    N = 2 * B * Cc
    dev = eeg1.device
    tdt = _TORCH_DT[code]
    frames = 1 + T // hop
    img = torch.empty(N, bins, frames, dtype=torch.float32, device=dev)
    TO.call("stft_logmag", eeg1, eeg2, window, img, B * Cc, T, n_fft,
           hop, bins)
    H1, W1, Wp, RP, slack = _spec_geometry(bins, frames)
    p1 = torch.empty((N * RP + slack) * 32, dtype=tdt, device=dev)
    # arg-max record of the fused max-pool (3 bits per pooled output) for the conv-1 weight gradient; int16 storage
    amax = torch.empty(N * H1 * W1 * 8, dtype=torch.int16, device=dev) if want_amax else None
    TO.call("spec_conv1_pool_fwd", img, w1, b1, p1, code, N, bins,
           frames, p1.numel(), amax)
    # every row the pooling reads (m + Wp + 1, m < N * RP) is written by the convolution: no memset
    y2 = torch.empty((N * RP + slack) * 64, dtype=tdt, device=dev)
    w2s = _spec_w2_seg(w2, code)
    if code == BF16 and _SPEC_DX_DIRECT and 128 + 2 * Wp + 2 <= 256:
        # implicit GEMM that stages each input tile once (nine row-shifted MMA operands per tile)
        TO.call("conv3x3_c32_c64", p1, p1.numel() // 32, w2s, b2.detach(), y2, N * RP, Wp + 1, Wp)
    else:
        a = TO.Operand(p1, 0, 0, 32, 0, 128, Wp)
        cm = _dense_matrix(TO.at(y2, (Wp + 1) * 64), code, 64)
        gemm(N * RP, 64, 384, code, a, TO.Operand(w2s, 0, 0, 384, 0, 0, 0), cm, bias=b2.detach())
    return img, p1, y2, (code, N, bins, frames, H1, W1, Wp, RP, slack, amax)


def _spec_padded_grad_buffer(meta, ch, device):
    """Uninitialised padded channels-last buffer whose slack rows (behind the last image) are zero: for kernels that write
    every position of every image themselves."""
    code, N, RP, slack = meta[0], meta[1], meta[7], meta[8]
    buf = torch.empty((N * RP + slack) * ch, dtype=_TORCH_DT[code], device=device)
    tail = buf[N * RP * ch:]
    TO.call("zero", tail, tail.numel() * tail.element_size())
    return buf


def _spec_front_bwd(dy2, img, p1, w1, b1, w2, meta, need_w1, need_w2, need_b2, db2=None):
    """Gradients of conv2 / conv1 given dY2 in the padded channels-last layout (db2: already computed by the producer)."""
    code, N, bins, frames, H1, W1, Wp, RP, slack, amax = meta
    dev, tdt = dy2.device, _TORCH_DT[code]
    dw1 = db1 = dw2 = None
    if need_b2 and db2 is None:
        db2 = torch.empty(64, dtype=torch.float32, device=dev)
        m = _dense_matrix(dy2, code, 64)
        TO.call("colsum", m, N * RP, 64, db2, 1)
    esz = dy2.element_size()
    if need_w2:
        # dW2[o, (kh, kw4, c)] = sum over flat padded positions of dY[m + Wp + 1, o] * P1[m + kh*Wp, kw4*32 + c]
        if code == BF16 and _SPEC_DX_DIRECT and 128 + 2 * Wp + 3 <= 256:
            # both operands staged once per 128-position tile; the four position slots of a kernel row are the M atoms of
            # one MMA operand (conv3x3.cu); result transposed: [(kh, kw4, c), o]
            dwt = zeros((384, 64), torch.float32, dev)
            TO.call("conv3x3_dw_c32_c64", p1, p1.numel() // 32, dy2, dy2.numel() // 64, dwt, N * RP, Wp + 1, Wp)
            dw2 = dwt.view(3, 4, 32, 64)[:, :3].permute(3, 2, 0, 1)
        else:
            dw = torch.empty(64, 384, dtype=torch.float32, device=dev)
            gemm(64, 384, N * RP, code, TO.Operand(TO.at(dy2, (Wp + 1) * 64), 1, 0, 64, 0, 0, 0),
                 TO.Operand(p1, 1, 0, 32, 0, 128, Wp), _dense_matrix(dw, F32, 384), accumulate=2)
            dw2 = dw.view(64, 3, 4, 32)[:, :, :3, :].permute(0, 3, 1, 2)
    if need_w1:
        # dP1 (padded layout) = full correlation of dY with the flipped kernel: same implicit GEMM, K = 3 x 256
        w2f = _spec_w2_flip(w2, code)
        dp1 = torch.empty((N * RP + slack) * 32, dtype=tdt, device=dev)
        if code == BF16 and _SPEC_DX_SLOTS == 4 and _SPEC_DX_DIRECT and 128 + 2 * Wp + 2 <= 256:
            # implicit GEMM that stages each input tile once and issues the nine taps as row-shifted MMA operands
            TO.call("conv3x3_c64_c32", dy2, dy2.numel() // 64, w2f, dp1, N * RP, Wp + 1, Wp)
        else:
            Kd = 3 * _SPEC_DX_SLOTS * 64
            a = TO.Operand(dy2, 0, 0, 64, 0, _SPEC_DX_SLOTS * 64, Wp)
            cm = _dense_matrix(TO.at(dp1, (Wp + 1) * 32), code, 32)
            gemm(N * RP, 32, Kd, code, a, TO.Operand(w2f, 0, 0, Kd, 0, 0, 0), cm)
        dwb = zeros((320,), torch.float32, dev)
        TO.call("spec_conv1_pool_bwd", img, w1, b1, dp1, code,
               dwb, dwb[288:], N, bins, frames, amax)
        dw1 = dwb[:288].view(32, 1, 3, 3)
        db1 = dwb[288:]
    return dw1, db1, dw2, db2


def _spec_padded_to_nchw(buf, meta, ch):
    """zero-bordered channels-last image buffer -> dense (N, ch, H1, W1) fp32."""
    code, N, bins, frames, H1, W1, Wp, RP, slack = meta[:9]
    out = torch.empty(N, ch, H1, W1, dtype=torch.float32, device=buf.device)
    copy_strided4(buf, out, (N, ch, H1, W1), (RP * ch, 1, Wp * ch, ch), (ch * H1 * W1, H1 * W1, W1, 1),
                  src_offset=(Wp + 1) * ch)
    return out


def _spec_nchw_to_padded(x, meta, ch):
    """dense (N, ch, H1, W1) -> zero-bordered channels-last buffer in the compute dtype."""
    code, N, bins, frames, H1, W1, Wp, RP, slack = meta[:9]
    x = x.contiguous()
    buf = zeros(((N * RP + slack) * ch,), _TORCH_DT[code], x.device)
    copy_strided4(x, buf, (N, ch, H1, W1), (ch * H1 * W1, H1 * W1, W1, 1), (RP * ch, 1, Wp * ch, ch),
                  dst_offset=(Wp + 1) * ch)
    return buf


class SpectrogramCNNFn(torch.autograd.Function):
    """(B,C,T) x2 -> pooled features [2B*C, 1024] (dual_eeg_transformer.py:98-127)."""

    @staticmethod
    def forward(ctx, eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins):
        _require_cuda(eeg1, eeg2, w1, w2)
        img, p1, y2, meta = _spec_front_fwd(eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins,
                                            want_amax=ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        code, N, bins, frames, H1, W1, Wp, RP, slack = meta[:9]
        pooled = torch.empty(N, 1024, dtype=_TORCH_DT[code], device=img.device)
        TO.call("relu_avgpool_fwd", y2, pooled, code, N, H1, W1)
        ctx.save_for_backward(img, p1, y2, w1, b1, w2)
        ctx.meta = meta
        return pooled

    @staticmethod
    def backward(ctx, dpool):
        code, N, bins, frames, H1, W1, Wp, RP, slack = ctx.meta[:9]
        img, p1, y2, w1, b1, w2 = ctx.saved_tensors
        dev, tdt = dpool.device, _TORCH_DT[code]
        dpool = dpool.contiguous()
        if _code(dpool) != code:
            dpool = cast(dpool, code)
        need = ctx.needs_input_grad
        dy2 = _spec_padded_grad_buffer(ctx.meta, 64, dev)
        db2 = zeros((64,), torch.float32, dev) if need[6] else None      # column sums of dy2, out of the same kernel
        TO.call("relu_avgpool_bwd", y2, dpool, dy2, code, N, H1, W1, db2)
        dw1, db1, dw2, db2 = _spec_front_bwd(dy2, img, p1, w1, b1, w2, ctx.meta, need[3] or need[4], need[5], need[6], db2)
        return None, None, None, dw1, db1, dw2, db2, None, None, None, None


def spectrogram_cnn(eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins):
    return SpectrogramCNNFn.apply(eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins)


# -- the same pipeline split at the spec_conv[3] output, for analysis hooks on that module (Grad-CAM,
#    5_Metrics/eeg_metrics.py:841): the kernels' own conv-2 output and its gradient are handed out as NCHW fp32 tensors
class SpecConvFrontFn(torch.autograd.Function):
    """(B,C,T) x2 -> (spec_conv[2] output (N,32,H1,W1), spec_conv[3] output (N,64,H1,W1)), fp32 NCHW, N = 2*B*C."""

    @staticmethod
    def forward(ctx, eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins):
        _require_cuda(eeg1, eeg2, w1, w2)
        img, p1, y2, meta = _spec_front_fwd(eeg1, eeg2, window, w1, b1, w2, b2, code, n_fft, hop, bins,
                                            want_amax=ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        ctx.save_for_backward(img, p1, w1, b1, w2)
        ctx.meta = meta
        p1n = _spec_padded_to_nchw(p1, meta, 32)
        ctx.mark_non_differentiable(p1n)
        return p1n, _spec_padded_to_nchw(y2, meta, 64)

    @staticmethod
    def backward(ctx, _dp1, dy2n):
        img, p1, w1, b1, w2 = ctx.saved_tensors
        need = ctx.needs_input_grad
        dy2 = _spec_nchw_to_padded(dy2n.float(), ctx.meta, 64)
        dw1, db1, dw2, db2 = _spec_front_bwd(dy2, img, p1, w1, b1, w2, ctx.meta, need[3] or need[4], need[5], need[6])
        return None, None, None, dw1, db1, dw2, db2, None, None, None, None


class SpecPoolFn(torch.autograd.Function):
    """spec_conv[3] output (N,64,H1,W1) fp32 -> ReLU -> AdaptiveAvgPool(4,4) -> flatten [N, 1024] (compute dtype)."""

    @staticmethod
    def forward(ctx, y2n, code, bins, frames):
        N = y2n.shape[0]
        H1, W1, Wp, RP, slack = _spec_geometry(bins, frames)
        meta = (code, N, bins, frames, H1, W1, Wp, RP, slack)
        y2 = _spec_nchw_to_padded(y2n.float(), meta, 64)
        pooled = torch.empty(N, 1024, dtype=_TORCH_DT[code], device=y2n.device)
        TO.call("relu_avgpool_fwd", y2, pooled, code, N, H1, W1)
        ctx.save_for_backward(y2)
        ctx.meta = meta
        return pooled

    @staticmethod
    def backward(ctx, dpool):
        code, N, bins, frames, H1, W1, Wp, RP, slack = ctx.meta[:9]
        (y2,) = ctx.saved_tensors
        dpool = dpool.contiguous()
        if _code(dpool) != code:
            dpool = cast(dpool, code)
        dy2 = _spec_padded_grad_buffer(ctx.meta, 64, dpool.device)
        TO.call("relu_avgpool_bwd", y2, dpool, dy2, code, N, H1, W1, None)
        return _spec_padded_to_nchw(dy2, ctx.meta, 64), None, None, None


def spectrogram_conv2_nchw(eeg1, eeg2, window, w1, b1, w2, b2, n_fft, hop, bins):
    """Analysis helper: the spec_conv[3] output (N, 64, H1, W1) in fp32, computed by the same kernels."""
    with torch.no_grad():
        img, p1, y2, meta = _spec_front_fwd(eeg1, eeg2, window, w1, b1, w2, b2, F32, n_fft, hop, bins)
        return _spec_padded_to_nchw(y2, meta, 64)


# ------------------------------------------------------------------------------------------------------
# IBS connectivity (parameter-free, no gradient) and the token-axis instance norm
# ------------------------------------------------------------------------------------------------------
_twiddles = {}


def _twiddle(T, device):
    key = (T, str(device))
    if key not in _twiddles:
        k = torch.arange(T // 2, dtype=torch.float64)
        ang = -2.0 * math.pi * k / T
        _twiddles[key] = torch.stack([torch.cos(ang), torch.sin(ang)], dim=1).to(torch.float32).to(device).contiguous()
    return _twiddles[key]


def band_bins(T, fs, bands):
    """Inclusive rfft-bin range of each band under the reference's mask (freqs >= lo) & (freqs <= hi)
    (dual_eeg_transformer.py:548-551), evaluated in fp32 like torch.fft.rfftfreq."""
    freqs = torch.fft.rfftfreq(T, d=1.0 / fs)
    lo_hi = []
    for lo, hi in bands:
        idx = torch.nonzero((freqs >= lo) & (freqs <= hi)).flatten()
        lo_hi.append((int(idx[0]), int(idx[-1])) if idx.numel() else (1, 0))
    return lo_hi


def ibs_connectivity(eeg1, eeg2, fs, bands, feature_indices, chunk=256):
    """(B,C,T) x2 fp32 -> (B, n_bands, len(feature_indices), C, C) fp32.  Batch is processed in chunks so the
    phase / band-passed scratch stays bounded."""
    _require_cuda(eeg1, eeg2)
    with torch.no_grad():
        eeg1 = eeg1.detach().contiguous().float()
        eeg2 = eeg2.detach().contiguous().float()
        B, Cc, T = eeg1.shape
        dev = eeg1.device
        nb, nf = len(bands), len(feature_indices)
        bins = band_bins(T, fs, bands)
        lo = [b[0] for b in bins]
        hi = [b[1] for b in bins]
        slot = [-1] * 7
        for s, f in enumerate(feature_indices):
            slot[f] = s
        slot_c = list(slot)
        valid = [b for b in bins if b[0] <= b[1]]
        nbins = (max(b[1] for b in valid) - min(b[0] for b in valid) + 1) if valid else 1
        out = torch.empty(B, nb, nf, Cc, Cc, dtype=torch.float32, device=dev)
        tw = _twiddle(T, dev)
        cb = min(B, chunk)
        phase = torch.empty(cb * nb * 2 * Cc * T, dtype=torch.float32, device=dev)
        xb = torch.empty_like(phase)
        stats = torch.empty(cb * nb * 2 * Cc * 8, dtype=torch.float32, device=dev)
        pspec = torch.empty(cb * 2 * Cc * nbins, dtype=torch.float32, device=dev)
        for b0 in range(0, B, cb):
            n = min(cb, B - b0)
            TO.call("ibs_connectivity", eeg1[b0:], eeg2[b0:], tw, phase,
                   xb, stats, pspec, out[b0:], n, Cc, T, nb, lo, hi, slot_c,
                   nf)
        return out


class InstNormTokensFn(torch.autograd.Function):
    """x (B, NT, P) fp32 -> y (compute dtype): per (b, p) normalisation over the NT tokens + affine (det:893-901)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, code, apply_norm):
        _require_cuda(x)
        x = x.contiguous().float()
        B, NT, P = x.shape
        y = torch.empty(B, NT, P, dtype=_TORCH_DT[code], device=x.device)
        TO.call("instnorm_tokens_fwd", x, _p(gamma), _p(beta), y, code, B, NT, P, 1e-5,
               1 if apply_norm else 0)
        ctx.save_for_backward(x)
        ctx.meta = (code, apply_norm)
        return y

    @staticmethod
    def backward(ctx, dy):
        code, apply_norm = ctx.meta
        if not apply_norm:
            return None, None, None, None, None
        (x,) = ctx.saved_tensors
        B, NT, P = x.shape
        dy = dy.contiguous()
        if _code(dy) != code:
            dy = cast(dy, code)
        dgb = zeros((2, P), torch.float32, x.device)
        TO.call("instnorm_tokens_bwd", x, dy, code, dgb[0], dgb[1], B, NT,
               P, 1e-5)
        return None, dgb[0], dgb[1], None, None


def instnorm_tokens(x, gamma, beta, code, apply_norm=True):
    return InstNormTokensFn.apply(x, gamma, beta, code, apply_norm)


# ------------------------------------------------------------------------------------------------------
# sequence assembly, pooling tail, concat
# ------------------------------------------------------------------------------------------------------
class SeqAssembleFn(torch.autograd.Function):
    """[cls | ibs | spec | h] + pos_embed[:L]  ->  X [2B, L, D]   (dual_eeg_transformer.py:1157-1179)."""

    @staticmethod
    def forward(ctx, cls, pos, ibs, spec, h, code):
        _require_cuda(h, cls, pos)
        S, n_h, D = h.shape
        B = S // 2
        n_ibs = ibs.shape[1] if ibs is not None else 0
        n_spec = spec.shape[1] if spec is not None else 0
        Lq = 1 + n_ibs + n_spec + n_h
        if Lq > pos.shape[0]:
            raise IndexError("index out of range in self: sequence length %d exceeds max_len %d" % (Lq, pos.shape[0]))
        ibs = cast(ibs.contiguous(), code) if ibs is not None else None
        spec = cast(spec.contiguous(), code) if spec is not None else None
        h = cast(h.contiguous(), code)
        out = torch.empty(S, Lq, D, dtype=_TORCH_DT[code], device=h.device)
        TO.call("seq_assemble_fwd", cls, pos, _p(ibs), _p(spec), h, out,
               code, S, B, Lq, D, n_ibs, n_spec, n_h)
        ctx.meta = (code, S, B, Lq, D, n_ibs, n_spec, n_h, pos.shape[0])
        return out

    @staticmethod
    def backward(ctx, dx):
        code, S, B, Lq, D, n_ibs, n_spec, n_h, max_len = ctx.meta
        dx = dx.contiguous()
        if _code(dx) != code:
            dx = cast(dx, code)
        dpos = zeros((max_len, D), torch.float32, dx.device)
        dibs = torch.empty(B, n_ibs, D, dtype=dx.dtype, device=dx.device) if n_ibs else None
        TO.call("seq_assemble_bwd", dx, dpos, _p(dibs), code, S, B, Lq, D, n_ibs)
        # cls also collects pos row 0's sum: both parameters receive the same batch-summed row, as SEPARATE tensors
        # (autograd keeps the returned tensors as .grad; aliased gradients would be scaled twice by in-place clipping /
        # GradScaler.unscale_)
        dcls = dpos[0].clone().view(1, 1, D)
        dspec = dx[:, 1 + n_ibs:1 + n_ibs + n_spec] if n_spec else None
        dh = dx[:, 1 + n_ibs + n_spec:]
        return dcls, dpos, dibs, dspec, dh, None


def seq_assemble(cls, pos, ibs, spec, h, code):
    return SeqAssembleFn.apply(cls, pos, ibs, spec, h, code)


class AddRowsBroadcastFn(torch.autograd.Function):
    """x [B, NT, D] + e [1, NT, D] (fp32 parameter) -- the IBS type embedding (det:909)."""

    @staticmethod
    def forward(ctx, x, e):
        x = x.contiguous()
        B, NT, D = x.shape
        out = torch.empty_like(x)
        TO.call("add_rows_broadcast", x, e, out, _code(x), B * NT, NT, D)
        ctx.meta = (B, NT, D)
        return out

    @staticmethod
    def backward(ctx, dy):
        B, NT, D = ctx.meta
        de = None
        if ctx.needs_input_grad[1]:
            dyc = dy.contiguous()
            de = zeros((NT, D), torch.float32, dy.device)
            TO.call("seq_assemble_bwd", dyc, de, None, _code(dyc), B, B, NT, D, 0)
            de = de.view(1, NT, D)
        return dy, de


def add_broadcast_rows(x, e):
    return AddRowsBroadcastFn.apply(x, e)


def ibs_scalar_features(eeg1, eeg2, fs, bands, chunk=256):
    """Legacy scalar IBS token features (det:418-470): (B,C,T) x2 fp32 -> (B, 7 * n_bands) fp32, no gradient."""
    _require_cuda(eeg1, eeg2)
    with torch.no_grad():
        eeg1 = eeg1.detach().contiguous().float()
        eeg2 = eeg2.detach().contiguous().float()
        B, Cc, T = eeg1.shape
        dev = eeg1.device
        nb = len(bands)
        bins = band_bins(T, fs, bands)
        lo = [b[0] for b in bins]
        hi = [b[1] for b in bins]
        valid = [b for b in bins if b[0] <= b[1]]
        nbins = (max(b[1] for b in valid) - min(b[0] for b in valid) + 1) if valid else 1
        out = torch.empty(B, nb * 7, dtype=torch.float32, device=dev)
        tw = _twiddle(T, dev)
        cb = min(B, chunk)
        phase = torch.empty(cb * nb * 2 * Cc * T, dtype=torch.float32, device=dev)
        xb = torch.empty_like(phase)
        stats = torch.empty(cb * nb * 2 * Cc * 8, dtype=torch.float32, device=dev)
        pspec = torch.empty(cb * 2 * Cc * nbins, dtype=torch.float32, device=dev)
        cspec = torch.empty(cb * 2 * Cc * nbins * 2, dtype=torch.float32, device=dev)
        for b0 in range(0, B, cb):
            n = min(cb, B - b0)
            TO.call("ibs_scalar_features", eeg1[b0:], eeg2[b0:], tw, phase,
                   xb, stats, pspec, cspec, out[b0:], n, Cc, T, nb,
                   lo, hi)
        return out


class TailPoolFn(torch.autograd.Function):
    """Z [2B, L, D] -> cls1, cls2 (B,D), sym (B,3D), mp (B,2D), ibs_pool (B,D); all fp32 (det:1193-1225)."""

    @staticmethod
    def forward(ctx, z, n_ibs, offset, ibs_single):
        _require_cuda(z)
        z = z.contiguous()
        S, Lq, D = z.shape
        B = S // 2
        dev = z.device
        f = torch.float32
        cls1 = torch.empty(B, D, dtype=f, device=dev)
        cls2 = torch.empty(B, D, dtype=f, device=dev)
        sym = torch.empty(B, 3 * D, dtype=f, device=dev)
        zf = torch.empty(B, 3 * D, dtype=f, device=dev)
        ibs_pool = torch.empty(B, D, dtype=f, device=dev) if n_ibs > 0 else None
        TO.call("tail_pool_fwd", z, _code(z), cls1, cls2, sym, zf,
               _p(ibs_pool), B, Lq, D, n_ibs, offset, 1 if ibs_single else 0)
        ctx.save_for_backward(z)
        ctx.meta = (n_ibs, offset, ibs_single)
        mp = zf[:, D:]
        if ibs_pool is None:
            return cls1, cls2, sym, mp
        return cls1, cls2, sym, mp, ibs_pool

    @staticmethod
    def backward(ctx, dcls1, dcls2, dsym, dmp, dibs=None):
        (z,) = ctx.saved_tensors
        n_ibs, offset, ibs_single = ctx.meta
        S, Lq, D = z.shape
        B = S // 2
        dev = z.device

        def f32c(t):
            return None if t is None else t.contiguous().float()
        dcls1, dcls2, dibs = f32c(dcls1), f32c(dcls2), f32c(dibs)
        dsym = f32c(dsym) if dsym is not None else zeros((B, 3 * D), torch.float32, dev)
        dzf = zeros((B, 3 * D), torch.float32, dev)
        if dmp is not None:
            copy_strided4(dmp.contiguous().float(), dzf, (1, 1, B, 2 * D), (0, 0, 2 * D, 1), (0, 0, 3 * D, 1), dst_offset=D)
        dz = torch.empty_like(z)
        TO.call("tail_pool_bwd", z, _code(z), _p(dcls1), _p(dcls2), dsym, dzf, _p(dibs),
               dz, B, Lq, D, n_ibs, offset, 1 if ibs_single else 0)
        return dz, None, None, None


def tail_pool(z, n_ibs, offset, ibs_single):
    return TailPoolFn.apply(z, n_ibs, offset, ibs_single)


class Concat2Fn(torch.autograd.Function):
    """[a | b] along the last dim of 2-D fp32 tensors (classifier input [f_pair, mp1, mp2], det:1212)."""

    @staticmethod
    def forward(ctx, a, b):
        B, Da = a.shape
        Db = b.shape[1]
        out = torch.empty(B, Da + Db, dtype=torch.float32, device=a.device)
        copy_strided4(a, out, (1, 1, B, Da), (0, 0, a.stride(0), a.stride(1)), (0, 0, Da + Db, 1))
        copy_strided4(b, out, (1, 1, B, Db), (0, 0, b.stride(0), b.stride(1)), (0, 0, Da + Db, 1), dst_offset=Da)
        ctx.meta = (Da, Db)
        return out

    @staticmethod
    def backward(ctx, d):
        Da, Db = ctx.meta
        return d[:, :Da], d[:, Da:]


def concat2(a, b):
    return Concat2Fn.apply(a, b)


# ------------------------------------------------------------------------------------------------------
# losses / fusion head
# ------------------------------------------------------------------------------------------------------
class CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        _require_cuda(logits, labels)
        logits = logits.contiguous().float()
        labels = labels.contiguous().long()
        B, Cn = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dl = torch.empty_like(logits)
        TO.call("cross_entropy", logits, labels, loss, dl, B, Cn)
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        g = g.contiguous().float()
        out = torch.empty_like(dl)
        TO.call("scale_by_device_scalar", dl, g, out, dl.numel())
        return out, None


def cross_entropy(logits, labels):
    return CrossEntropyFn.apply(logits, labels)


# ------------------------------------------------------------------------------------------------------
# batch-level auxiliary losses (dual_eeg_transformer.py:1255-1371); fp32 (B, d) tokens
# ------------------------------------------------------------------------------------------------------
class L2NormalizeRowsFn(torch.autograd.Function):
    """F.normalize(x, p=2, dim=-1) on a 2-D fp32 tensor."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        x = x.contiguous().float()
        R, D = x.shape
        y = torch.empty_like(x)
        inv = torch.empty(R, dtype=torch.float32, device=x.device)
        TO.call("l2norm_rows_fwd", x, y, inv, R, D, 1e-12)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        TO.call("l2norm_rows_bwd", dy, y, inv, dx, y.shape[0], y.shape[1])
        return dx


def l2_normalize_rows(x):
    return L2NormalizeRowsFn.apply(x)


class StackRowsFn(torch.autograd.Function):
    """torch.cat([a, b], dim=0) of 2-D fp32 tensors."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous().float(), b.contiguous().float()
        out = torch.empty(a.shape[0] + b.shape[0], a.shape[1], dtype=torch.float32, device=a.device)
        D = a.shape[1]
        copy_strided4(a, out, (1, 1, a.shape[0], D), (0, 0, D, 1), (0, 0, D, 1))
        copy_strided4(b, out, (1, 1, b.shape[0], D), (0, 0, D, 1), (0, 0, D, 1), dst_offset=a.shape[0] * D)
        ctx.na = a.shape[0]
        return out

    @staticmethod
    def backward(ctx, d):
        return d[:ctx.na], d[ctx.na:]


class SimilarityLossFn(torch.autograd.Function):
    """loss(a . b^T / temperature) for unit-norm rows a [B, d], b [N, d] (fp32):
    kind 'infonce' = cross_entropy(sim, arange(B)) (det:1290-1302);  kind 'supcon' = supervised contrastive loss over
    sim = a . a^T with labels (det:1336-1371; b is a).  The row kernel leaves d loss / d sim in place of sim, so the
    backward pass is the two GEMMs  da = G b / T,  db = G^T a / T."""

    @staticmethod
    def forward(ctx, a, b, labels, temperature, kind):
        _require_cuda(a, b)
        a = a.contiguous().float()
        same = b is None
        b = a if same else b.contiguous().float()
        B, D = a.shape
        N = b.shape[0]
        dev = a.device
        sim = torch.empty(B, N, dtype=torch.float32, device=dev)
        gemm(B, N, D, F32, TO.Operand(a, 0, 0, D, 0, 0, 0), TO.Operand(b, 0, 0, D, 0, 0, 0),
             _dense_matrix(sim, F32, N), alpha=1.0 / temperature)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        if kind == "infonce":
            TO.call("infonce_rows", sim, loss, B, N)
        else:
            labels = labels.contiguous().long()
            scratch = torch.empty(3 * B + 2, dtype=torch.float32, device=dev)
            TO.call("supcon_rows", sim, labels, scratch, scratch[3 * B:],
                   loss, B)
        ctx.save_for_backward(a, b, sim)
        ctx.meta = (temperature, same)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b, G = ctx.saved_tensors
        temperature, same = ctx.meta
        B, D = a.shape
        N = b.shape[0]
        dev = a.device
        g = g.contiguous().float()
        # da[B, d] = G[B, N] . b[N, d]   (b consumed as an MN-major operand);   db[N, d] = G^T . a
        da = torch.empty(B, D, dtype=torch.float32, device=dev)
        gemm(B, D, N, F32, TO.Operand(G, 0, 0, N, 0, 0, 0), TO.Operand(b, 1, 0, D, 0, 0, 0),
             _dense_matrix(da, F32, D), alpha=1.0 / temperature)
        db = torch.empty(N, D, dtype=torch.float32, device=dev)
        gemm(N, D, B, F32, TO.Operand(G, 1, 0, N, 0, 0, 0), TO.Operand(a, 1, 0, D, 0, 0, 0),
             _dense_matrix(db, F32, D), alpha=1.0 / temperature)
        if same:
            da = da + db          # sim = a a^T: both factors are the same tensor
            db = None
        out_a = torch.empty_like(da)
        TO.call("scale_by_device_scalar", da, g, out_a, da.numel())
        out_b = None
        if db is not None:
            out_b = torch.empty_like(db)
            TO.call("scale_by_device_scalar", db, g, out_b, db.numel())
        return out_a, out_b, None, None, None


def infonce_loss(a, b, temperature):
    return SimilarityLossFn.apply(a, b, None, float(temperature), "infonce")


def supcon_loss(a, labels, temperature):
    return SimilarityLossFn.apply(a, None, labels, float(temperature), "supcon")


class MseLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        a, b = a.contiguous().float(), b.contiguous().float()
        da = torch.empty_like(a)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        TO.call("mse_loss", a, b, da, loss, a.numel())
        ctx.save_for_backward(da)
        return loss

    @staticmethod
    def backward(ctx, g):
        (da,) = ctx.saved_tensors
        out = torch.empty_like(da)
        TO.call("scale_by_device_scalar", da, g.contiguous().float(), out, da.numel())
        return out, -out


def mse_loss(a, b):
    return MseLossFn.apply(a, b)


FUZZY_MODES = {"full": 0, "no_temperature": 1, "no_fuzzification": 2, "fixed_weights": 3}


def _fuzzy_desc(params, mode, B, Cn, eps_temp, eps_log, eps_div):
    d = TO.FuzzyDesc()
    (d.tau_img, d.tau_eeg, d.c_reliable, d.c_unreliable_img, d.c_unreliable_eeg, d.log_sigma_reliable_img,
     d.log_sigma_reliable_eeg, d.log_sigma_unreliable_img, d.log_sigma_unreliable_eeg, d.beta) = list(params)
    d.mode, d.B, d.num_classes = mode, B, Cn
    d.eps_temp, d.eps_log, d.eps_div = eps_temp, eps_log, eps_div
    return d


class FuzzyGatingFn(torch.autograd.Function):
    """params order: tau_img, tau_eeg, c_reliable(buffer), c_unreliable_img, c_unreliable_eeg, log_sigma_reliable_img,
    log_sigma_reliable_eeg, log_sigma_unreliable_img, log_sigma_unreliable_eeg, beta."""

    @staticmethod
    def forward(ctx, img, eeg, mode, eps_temp, eps_log, eps_div, *params):
        _require_cuda(img, eeg, *params)
        img = img.contiguous().float()
        eeg = eeg.contiguous().float()
        B, Cn = img.shape
        dev = img.device
        fused = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        alpha = torch.empty(B, dtype=torch.float32, device=dev)
        aux = torch.empty(B + 1, 16, dtype=torch.float32, device=dev)
        d = _fuzzy_desc(params, mode, B, Cn, eps_temp, eps_log, eps_div)
        TO.call("fuzzy_fwd", d, img, eeg, fused, alpha,
               aux)
        ctx.save_for_backward(img, eeg, *params)
        ctx.meta = (mode, eps_temp, eps_log, eps_div)
        ctx.mark_non_differentiable(aux)
        return fused, alpha, aux

    @staticmethod
    def backward(ctx, g_fused, g_alpha, _g_aux):
        mode, eps_temp, eps_log, eps_div = ctx.meta
        saved = ctx.saved_tensors
        img, eeg, params = saved[0], saved[1], saved[2:]
        B, Cn = img.shape
        dev = img.device
        g_fused = g_fused.contiguous().float() if g_fused is not None else zeros((B, Cn), torch.float32, dev)
        g_alpha = g_alpha.contiguous().float() if g_alpha is not None else None
        d_img = torch.empty_like(img)
        d_eeg = torch.empty_like(eeg)
        dp = zeros((12,), torch.float32, dev)
        d = _fuzzy_desc(params, mode, B, Cn, eps_temp, eps_log, eps_div)
        TO.call("fuzzy_bwd", d, img, eeg, g_fused, _p(g_alpha), d_img,
               d_eeg, dp)
        grads = (dp[0], dp[1], None, dp[2], dp[3], dp[4], dp[5], dp[6], dp[7], dp[8:12])
        return (d_img, d_eeg, None, None, None, None) + grads


def fuzzy_gating(img, eeg, mode, eps_temp, eps_log, eps_div, params):
    return FuzzyGatingFn.apply(img, eeg, mode, float(eps_temp), float(eps_log), float(eps_div), *params)


# ------------------------------------------------------------------------------------------------------
# ViT patch embedding (+ input fusion + cls + pos)
# ------------------------------------------------------------------------------------------------------
PATCH_MODES = {"concat": 0, "add": 1, "subtract": 2, "subtract_abs": 3, "multiply": 4, "single": 5, "single_pair": 6}


def _img_view(t):
    """(tensor, batch_stride) of a (B,3,H,W) fp32 image batch; channel-sliced views of a wider tensor are legal."""
    t = t.float()
    B, Cc, H, W = t.shape
    if not (t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W and t.stride(0) % 4 == 0):
        t = t.contiguous()
    return t, t.stride(0)


class VitEmbedFn(torch.autograd.Function):
    """Fuse -> patchify -> patch GEMM (+bias +pos) -> cls row.  img_a, img_b: (B,3,H,W) fp32.
    mode 0..4: EarlyFusionViT input fusion (one token sequence per trial);
    mode 5   : single image (img_b None) ;  mode 6: two single images stacked along the batch (LateFusionViT)."""

    @staticmethod
    def forward(ctx, img_a, img_b, w, b, cls, pos, mode, code, ps):
        _require_cuda(img_a, w)
        img_a, a_bs = _img_view(img_a)
        if img_b is not None:
            img_b, b_bs = _img_view(img_b)
        else:
            img_b, b_bs = img_a, a_bs
        Bi, _, H, W = img_a.shape
        D = w.shape[0]
        n = (H // ps) * (W // ps)
        cin = 6 if mode == 0 else 3
        K = cin * ps * ps
        if w.shape[1] != cin:
            raise RuntimeError("patch embedding expects %d input channels, weight has %d" % (cin, w.shape[1]))
        if pos.shape[1] != n + 1:
            raise RuntimeError("pos_embed has %d tokens, image gives %d" % (pos.shape[1], n + 1))
        dev, tdt = img_a.device, _TORCH_DT[code]
        B = 2 * Bi if mode == 6 else Bi
        patches = torch.empty(B * n, K, dtype=tdt, device=dev)
        stats = torch.empty(Bi * 6, dtype=torch.float32, device=dev) if mode == 4 else None
        if mode == 6:
            TO.call("vit_patchify", img_a, img_a, a_bs, a_bs, patches, None, code, Bi,
                   H, W, ps, 5)
            TO.call("vit_patchify", img_b, img_b, b_bs, b_bs,
                   TO.at(patches, Bi * n * K), None, code, Bi, H, W, ps, 5)
        else:
            TO.call("vit_patchify", img_a, img_b, a_bs, b_bs, patches, _p(stats), code,
                   Bi, H, W, ps, mode)
        w2 = weight_plain(w, code)
        out = torch.empty(B, n + 1, D, dtype=tdt, device=dev)
        pos2 = pos.detach().reshape(n + 1, D)
        cm = TO.Matrix(TO.at(out, D), code, n, D, (n + 1) * D)
        if code == BF16:
            # the position table in the compute dtype (persistent copy, refreshed with the weights): with an fp32 residual the
            # GEMM fell back to the generic epilogue (221 us at 0.39 of the tensor peak for the ViT-B patch embedding)
            rm = TO.Matrix(TO.at(weight_plain(pos, code).reshape(n + 1, D), D), code, n, D, 0)
        else:
            rm = TO.Matrix(TO.at(pos2, D), F32, n, D, 0)
        gemm(B * n, D, K, code, TO.Operand(patches, 0, 0, K, 0, 0, 0), TO.Operand(w2, 0, 0, K, 0, 0, 0),
             cm, bias=b.detach() if b is not None else None, residual=rm)
        TO.call("fill_row0", cls, pos2, out, code, B, n + 1, D)
        ctx.save_for_backward(patches, w)
        ctx.meta = (code, B, n, D, K, b is not None)
        return out

    @staticmethod
    def backward(ctx, dx):
        code, B, n, D, K, has_bias = ctx.meta
        patches, w = ctx.saved_tensors
        dev = dx.device
        dx = dx.contiguous()
        if _code(dx) != code:
            dx = cast(dx, code)
        need = ctx.needs_input_grad
        dw = db = dcls = dpos = None
        if need[2] or need[3]:
            dtok = torch.empty(B * n, D, dtype=dx.dtype, device=dev)   # dense copy of dx[:, 1:, :]
            copy_strided4(dx, dtok, (1, B, n, D), (0, (n + 1) * D, D, 1), (0, n * D, D, 1), src_offset=D)
            if need[2]:
                dw = _grad_weight(dtok, patches, D, K).view(w.shape)
            if need[3] and has_bias:
                db = colsum(dtok, D)
        if need[4] or need[5]:
            dp = zeros((n + 1, D), torch.float32, dev)
            TO.call("seq_assemble_bwd", dx, dp, None, code, B, B, n + 1, D, 0)
            dpos = dp.view(1, n + 1, D)
            dcls = dp[0].clone().view(1, 1, D)         # never alias two parameters' gradients (see SeqAssembleFn)
        return None, None, dw, db, dcls, dpos, None, None, None


def vit_embed(img_a, img_b, w, b, cls, pos, mode, code, ps=16):
    return VitEmbedFn.apply(img_a, img_b, w, b, cls, pos, mode, code, ps)
