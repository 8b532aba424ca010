"""ctypes binding of libtg_b200.so.  Prototypes are derived from include/tg_b200.h so that the header is the single
source of truth for the C ABI.  There is NO fallback: if the library is missing, or CUDA is unavailable when a kernel
is requested, an exception is raised (BASELINE.json north_star: "no CPU fallback")."""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
HEADER = os.path.join(REPO, 'include', 'tg_b200.h')
LIB_PATH = os.path.join(HERE, 'libtg_b200.so')


class TgError(RuntimeError):
    pass


class ConvGemm(ctypes.Structure):
    """tg_conv_gemm_t (include/tg_b200.h)."""
    _fields_ = [('A', ctypes.c_void_p), ('lda', ctypes.c_int), ('asc', ctypes.c_int), ('a_bstride', ctypes.c_longlong),
                ('W', ctypes.c_void_p), ('ldw', ctypes.c_int), ('wsj', ctypes.c_int), ('wsc', ctypes.c_int),
                ('Y', ctypes.c_void_p), ('ldc', ctypes.c_int),
                ('B', ctypes.c_int), ('Tout', ctypes.c_int), ('Tin', ctypes.c_int), ('N', ctypes.c_int), ('Cin', ctypes.c_int),
                ('taps', ctypes.c_int), ('stride', ctypes.c_int), ('dil', ctypes.c_int), ('pad', ctypes.c_int),
                ('ToutFull', ctypes.c_int), ('ostride', ctypes.c_int), ('ooff', ctypes.c_int),
                ('pscale', ctypes.c_void_p), ('pshift', ctypes.c_void_p), ('pslope', ctypes.c_float),
                ('escale', ctypes.c_void_p), ('bias', ctypes.c_void_p),
                ('act1', ctypes.c_int), ('slope1', ctypes.c_float),
                ('mask', ctypes.c_void_p), ('ldmask', ctypes.c_int),
                ('residual', ctypes.c_void_p), ('ldres', ctypes.c_int),
                ('act2', ctypes.c_int), ('accumulate', ctypes.c_int)]


class ConvWgrad(ctypes.Structure):
    """tg_conv_wgrad_t (include/tg_b200.h)."""
    _fields_ = [('A', ctypes.c_void_p), ('lda', ctypes.c_int),
                ('G', ctypes.c_void_p), ('ldg', ctypes.c_int),
                ('dW', ctypes.c_void_p), ('ldw', ctypes.c_int), ('wsj', ctypes.c_int), ('wsc', ctypes.c_int),
                ('B', ctypes.c_int), ('Tout', ctypes.c_int), ('Tin', ctypes.c_int), ('N', ctypes.c_int), ('Cin', ctypes.c_int),
                ('taps', ctypes.c_int), ('stride', ctypes.c_int), ('dil', ctypes.c_int), ('pad', ctypes.c_int),
                ('pscale', ctypes.c_void_p), ('pshift', ctypes.c_void_p), ('pslope', ctypes.c_float),
                ('dbias', ctypes.c_void_p)]


class GemmTf32(ctypes.Structure):
    """tg_gemm_tf32_t (include/tg_b200.h)."""
    _fields_ = [('A', ctypes.c_void_p), ('lda', ctypes.c_int), ('a_rows', ctypes.c_longlong),
                ('Bw', ctypes.c_void_p), ('ldb', ctypes.c_int),
                ('C', ctypes.c_void_p), ('ldc', ctypes.c_int),
                ('M', ctypes.c_int), ('N', ctypes.c_int), ('K', ctypes.c_int), ('taps', ctypes.c_int), ('shift0', ctypes.c_int),
                ('T', ctypes.c_int),
                ('escale', ctypes.c_void_p), ('bias', ctypes.c_void_p),
                ('act1', ctypes.c_int), ('slope1', ctypes.c_float),
                ('mask', ctypes.c_void_p), ('ldmask', ctypes.c_int),
                ('residual', ctypes.c_void_p), ('ldres', ctypes.c_int),
                ('act2', ctypes.c_int), ('accumulate', ctypes.c_int),
                ('clip_rows', ctypes.c_int), ('a_clip_pitch', ctypes.c_longlong)]


class WgradTf32(ctypes.Structure):
    """tg_wgrad_tf32_t (include/tg_b200.h)."""
    _fields_ = [('G', ctypes.c_void_p), ('ldg', ctypes.c_int), ('X', ctypes.c_void_p), ('ldx', ctypes.c_int),
                ('dW', ctypes.c_void_p), ('ldw', ctypes.c_int), ('dbias', ctypes.c_void_p),
                ('B', ctypes.c_int), ('T', ctypes.c_int), ('N', ctypes.c_int), ('Cin', ctypes.c_int), ('shift', ctypes.c_int),
                ('x_clip_pitch', ctypes.c_longlong)]


def _ctype(decl: str):
    d = decl.strip()
    if '*' in d:
        return ctypes.c_void_p
    if d.startswith('tg_stream'):
        return ctypes.c_void_p
    if 'unsigned long long' in d:
        return ctypes.c_ulonglong
    if 'long long' in d:
        return ctypes.c_longlong
    if d.startswith('size_t'):
        return ctypes.c_size_t
    if d.startswith('double'):
        return ctypes.c_double
    if d.startswith('float'):
        return ctypes.c_float
    if d.startswith('int'):
        return ctypes.c_int
    raise ValueError('unmapped C type in header: %r' % decl)


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{symbol: (restype, [argtypes])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'typedef\s+struct\s*\{.*?\}\s*\w+\s*;', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const\s+char\s*\*|int|size_t)\s+(tg_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if 'char' in ret else (ctypes.c_size_t if ret == 'size_t' else ctypes.c_int)
        argtypes = [] if args in ('', 'void') else [_ctype(a) for a in args.split(',')]
        protos[name] = (restype, argtypes)
    return protos


_lib = None
TRACE_ONLY = os.environ.get('TGB200_TRACE_ONLY', '') == '1'
trace = []          # (symbol,) records of the launch plan when TGB200_TRACE_ONLY=1


class _TraceLib:
    """Launch-plan tracer for CPU-only host-logic tests (TGB200_TRACE_ONLY=1): records which C-ABI entry points the
    host code WOULD call, executes nothing and produces no results.  It is not a fallback - outputs stay uninitialised."""

    def __init__(self, protos):
        self._protos = protos

    def __getattr__(self, name):
        if name not in self._protos:
            raise AttributeError(name)

        def fn(*args):
            if len(args) != len(self._protos[name][1]):
                raise TypeError('%s: expected %d arguments, got %d' % (name, len(self._protos[name][1]), len(args)))
            trace.append(name)
            if name in ('tg_gru_sync_ints', 'tg_gru_tf32_sync_ints'):
                return 64
            if name in ('tg_gru_bwd_scratch_floats', 'tg_gru_bwd_tf32_scratch_floats'):
                return 2 * 2 * args[0] * 8 * ((args[1] + 3) // 4 * 4)
            return 0
        return fn


def load() -> ctypes.CDLL:
    """Loads the shared library (no CUDA call is made) and installs the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if TRACE_ONLY:
        _lib = _TraceLib(parse_header())
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TgError('libtg_b200.so is not built: run `python __graft_entry__.py build` (needs nvcc). '
                      'There is no CPU / PyTorch fallback for the kernels.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError here == header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        raise TgError('%s failed (rc=%d): %s' % (what, rc, load().tg_last_error().decode()))


def require_cuda():
    import torch
    if TRACE_ONLY:
        return
    if not torch.cuda.is_available():
        raise TgError('tgb200 kernels need a CUDA device (sm_100a); there is no CPU fallback')
