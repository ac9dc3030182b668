"""``torch.library`` registration of the C ABI: every entry point of ``libeyegaze_b200.so`` (include/eyegaze_b200.h) is a
custom op ``torch.ops.eyegaze_b200.<name>`` that takes TENSORS where the C function takes device pointers.

This is the "thin C-ABI torch.library extension" of the north star: the host side (ops.py, optim.py, inputs.py -- the
autograd Functions and nn.Modules that mirror the reference's interface) never touches a raw pointer; it calls
``TO.call("layernorm_fwd", x, gamma, beta, y, mean, rstd, code, M, D, eps)`` and the registered CUDA implementation reads
``data_ptr()`` / the current stream (``torch.cuda.current_stream()``, i.e. at::cuda::getCurrentCUDAStream) and makes the C
call; a C-side failure surfaces as ``RuntimeError`` (SURVEY 8b error convention).

The schemas are generated from the ctypes prototypes in ``_lib._SIGNATURES`` -- one source of truth with the header:

  ``void* p``                  -> ``Tensor(x!)? p``   (optional; conservatively declared mutable -- these are out-variants)
  ``int32 / int64 / uint64``   -> ``int``
  ``float``                    -> ``float``
  ``const int32* / int64* / float*`` host arrays -> ``int[]`` / ``float[]``
  ``const float* const*`` (array of device scalars) -> ``Tensor[]``
  ``const egb_*_desc*``        -> the descriptor's fields, flattened in declaration order (nested operand / matrix structs
                                 included); build one with ``GemmDesc`` / ``AttentionDesc`` / ... of this module, which hold tensors
  trailing ``void* stream``    -> dropped (current stream)

Every op returns ``()`` and has a no-op Meta implementation, so programs can be traced / fake-tensor-propagated by name.
Ops without tensor arguments are registered for CompositeExplicitAutograd.  Host-only functions of the library (error
string, launch counter, profiling switches, seed-epoch creation) are not ops and stay on ``_lib``.
"""
import ctypes as C
from typing import Any, List, Sequence

import torch

from . import _lib as L

NAMESPACE = "eyegaze_b200"
_LIB = torch.library.Library(NAMESPACE, "DEF")
_HOST_ONLY = {"egb_prof_enable", "egb_prof_read", "egb_prof_dump", "egb_debug_attention_timing", "egb_debug_gemm_timing",
              "egb_seed_epoch_enable"}
_ARRAY_ELEM = {C.POINTER(L.i32): L.i32, C.POINTER(L.i64): L.i64, C.POINTER(L.f32): L.f32}
_INTS = (L.i32, L.i64, L.u64)
_LETTERS = "abcdefghijklmnopqrstuvwxyz"


def _is_struct_ptr(t) -> bool:
    return hasattr(t, "_type_") and isinstance(t._type_, type) and issubclass(t._type_, C.Structure)


def _flatten_fields(cls, prefix=""):
    """[(flat name, ctype)] of a ctypes Structure, nested structures expanded in declaration order."""
    out = []
    for name, ct in cls._fields_:
        if isinstance(ct, type) and issubclass(ct, C.Structure):
            out += _flatten_fields(ct, prefix + name + "_")
        else:
            out.append((prefix + name, ct))
    return out


class _TensorStruct:
    """Python twin of a ctypes descriptor that is filled field by field (attention / fuzzy / AdamW state): same field
    names, device-pointer fields hold tensors (or None)."""
    _names = ()
    _defaults = ()

    def __init__(self):
        for n, v in zip(self._names, self._defaults):
            setattr(self, n, v)

    def _flat(self):
        return [getattr(self, n) for n in self._names]


def _twin(ctype):
    flat = _flatten_fields(ctype)
    return type(ctype.__name__, (_TensorStruct,), {"_names": tuple(n for n, _ in flat),
                                                   "_defaults": tuple(None if ct is L.vp else 0 for _, ct in flat)})


AttentionDesc, FuzzyDesc, AdamwState = _twin(L.AttentionDesc), _twin(L.FuzzyDesc), _twin(L.AdamwState)


# The GEMM descriptor is built ~250 times per training step: its twins are plain tuples in flat field order.
def Operand(t, major, rows_per_group, row_stride, group_stride, seg_len=0, seg_row_shift=0):
    return (t, major, rows_per_group, row_stride, group_stride, seg_len, seg_row_shift)


def Matrix(t, dtype, rows_per_group, row_stride, group_stride):
    return (t, dtype, rows_per_group, row_stride, group_stride)


def GemmDesc(M, N, K, in_dtype, a, b, c, c_pre, residual, aux, bias, alpha, act, act_bwd, aux_scale, dropout_p,
             dropout_seed, accumulate, split_k, c_colsum):
    return (M, N, K, in_dtype) + a + b + c + c_pre + residual + aux + (bias, alpha, act, act_bwd, aux_scale, dropout_p,
                                                                      dropout_seed, accumulate, split_k, c_colsum)


_PLANS = {}
_GENERATED = {}


def _struct_expr(ctype, names, pos):
    """Python source of a positional ctypes constructor call for `ctype`, consuming argument names from `pos[0]` on."""
    parts = []
    for _fname, ct in ctype._fields_:
        if isinstance(ct, type) and issubclass(ct, C.Structure):
            parts.append(_struct_expr(ct, names, pos))
        else:
            a = names[pos[0]]
            pos[0] += 1
            if ct is L.vp:
                parts.append("(None if %s is None else %s.data_ptr())" % (a, a))
            elif ct is L.u64:
                parts.append("(%s & 0xFFFFFFFFFFFFFFFF)" % a)
            else:
                parts.append(a)
    return "_ct_%s(%s)" % (ctype.__name__, ", ".join(parts))


def _register(cname: str, argtypes: Sequence) -> None:
    """Defines eyegaze_b200::<name> and GENERATES its CUDA implementation as straight-line Python (no per-call loops over
    the argument plan: a GEMM descriptor has 48 fields and is issued ~250 times per step)."""
    name = cname[4:]
    has_stream = len(argtypes) > 0 and argtypes[-1] is L.vp
    # (every op of the library takes its stream last; the prototypes without one are the host-only functions above)
    cargs = list(argtypes[:-1]) if has_stream else list(argtypes)
    schema, params, cexprs, pre, env = [], [], [], [], {"_call": L.call, "_byref": C.byref, "_cname": cname,
                                                        "_cur": torch.cuda.current_stream}
    n_t = 0

    def tensor_arg(label):
        nonlocal n_t
        ann = _LETTERS[n_t % 26] * (1 + n_t // 26)
        n_t += 1
        return "Tensor(%s!)? %s" % (ann, label)
    for i, ct in enumerate(cargs):
        label = "a%d" % i
        if ct is L.vp:
            schema.append(tensor_arg(label))
            params.append(label)
            cexprs.append("(None if %s is None else %s.data_ptr())" % (label, label))
        elif ct in _INTS:
            schema.append("int " + label)
            params.append(label)
            cexprs.append("(%s & 0xFFFFFFFFFFFFFFFF)" % label if ct is L.u64 else label)
        elif ct is L.f32:
            schema.append("float " + label)
            params.append(label)
            cexprs.append(label)
        elif ct in _ARRAY_ELEM:
            elem = _ARRAY_ELEM[ct]
            schema.append(("float[] " if elem is L.f32 else "int[] ") + label)
            params.append(label)
            env["_arr%d" % i] = elem
            pre.append("    k%d = (_arr%d * len(%s))(*%s)" % (i, i, label, label))
            cexprs.append("k%d" % i)
        elif ct == C.POINTER(L.vp):
            schema.append("Tensor[] " + label)
            params.append(label)
            env["_vp"] = L.vp
            pre.append("    k%d = (_vp * len(%s))(*[t.data_ptr() for t in %s])" % (i, label, label))
            cexprs.append("k%d" % i)
        elif _is_struct_ptr(ct):
            flat = _flatten_fields(ct._type_)
            names = ["%s_%s" % (label, fname) for fname, _ in flat]
            for fname, fct in flat:
                full = "%s_%s" % (label, fname)
                if fct is L.vp:
                    schema.append(tensor_arg(full))
                elif fct in _INTS:
                    schema.append("int " + full)
                else:
                    schema.append("float " + full)
            params += names

            def collect(c):
                env["_ct_" + c.__name__] = c
                for _n, f in c._fields_:
                    if isinstance(f, type) and issubclass(f, C.Structure):
                        collect(f)
            collect(ct._type_)
            pre.append("    k%d = %s" % (i, _struct_expr(ct._type_, names, [0])))
            cexprs.append("_byref(k%d)" % i)
        else:
            raise TypeError("%s: no torch.library mapping for argument %d (%r)" % (cname, i, ct))
    if has_stream:
        cexprs.append("_cur().cuda_stream")
    src = "def impl(%s):\n%s\n    _call(_cname, %s)\n" % (", ".join(params), "\n".join(pre) if pre else "    pass",
                                                          ", ".join(cexprs))
    exec(src, env)
    _PLANS[name] = len(params)
    _GENERATED[name] = src
    _LIB.define("%s(%s) -> ()" % (name, ", ".join(schema)))
    _LIB.impl(name, env["impl"], "CUDA" if n_t else "CompositeExplicitAutograd")
    if n_t:
        _LIB.impl(name, lambda *a: None, "Meta")


for _cname, _argtypes in L._SIGNATURES.items():
    if _cname not in _HOST_ONLY:
        _register(_cname, _argtypes)

OP_NAMES = sorted(_PLANS)
_OPS = getattr(torch.ops, NAMESPACE)
_RESOLVED = {}


def call(name: str, *args) -> None:
    """``torch.ops.eyegaze_b200.<name>(...)``; a descriptor twin (tuple or field object) is splatted into its fields."""
    fn = _RESOLVED.get(name)
    if fn is None:
        fn = _RESOLVED[name] = getattr(_OPS, name).default
    if len(args) != _PLANS[name]:
        flat = []
        for a in args:
            if type(a) is tuple:
                flat += a
            elif isinstance(a, _TensorStruct):
                flat += a._flat()
            else:
                flat.append(a)
        args = flat
    fn(*args)


def at(t: torch.Tensor, elem_offset: int = 0) -> torch.Tensor:
    """1-D view of a contiguous buffer starting ``elem_offset`` elements in: the tensor form of ``ptr + offset``."""
    v = t.view(-1) if t.is_contiguous() else t.reshape(-1)
    return v[elem_offset:] if elem_offset else v
