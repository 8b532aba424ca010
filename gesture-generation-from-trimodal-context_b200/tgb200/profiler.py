"""Per-launch device timing of the C-ABI calls (CUDA events on the launching stream).  Used by bench.py to find the
dominant kernel of a step and its achieved FLOP/s or GB/s; never enabled inside a timed throughput region."""
from __future__ import annotations

import collections

import torch

from . import _lib


def _tag(name, args):
    try:
        if name in ('tg_conv_gemm_f32', 'tg_conv_wgrad_f32'):
            g = args[0]._obj
            return 'B%d Tout%d N%d Cin%d taps%d s%d' % (g.B, g.Tout, g.N, g.Cin, g.taps, g.stride)
        if name == 'tg_gemm_tf32':
            g = args[0]._obj
            return 'M%d N%d K%d taps%d' % (g.M, g.N, g.K, g.taps)
        if name == 'tg_wgrad_tf32':
            g = args[0]._obj
            return 'B%d T%d N%d Cin%d shift%d' % (g.B, g.T, g.N, g.Cin, g.shift)
        if name.startswith('tg_gru_layer'):
            return 'B%d T%d H%d' % (args[-4], args[-3], args[-2])
    except Exception:
        pass
    return ''


def _work(name, args):
    """(flops, bytes) of one call, from its arguments (algorithmic: 2*MACs; bytes = operands touched once)."""
    try:
        if name == 'tg_conv_gemm_f32':
            g = args[0]._obj
            m = g.B * g.Tout
            return 2.0 * m * g.N * g.Cin * g.taps, 4.0 * (m * g.Cin * min(g.taps, max(g.stride, 1)) + g.N * g.Cin * g.taps + m * g.N)
        if name == 'tg_conv_wgrad_f32':
            g = args[0]._obj
            m = g.B * g.Tout
            return 2.0 * m * g.N * g.Cin * g.taps, 4.0 * (m * g.Cin + m * g.N + g.N * g.Cin * g.taps)
        if name == 'tg_gemm_tf32':
            g = args[0]._obj
            return 2.0 * g.M * g.N * g.K * g.taps, 4.0 * (g.M * g.K + g.N * g.K * g.taps + g.M * g.N)
        if name == 'tg_wgrad_tf32':
            g = args[0]._obj
            return 2.0 * g.B * g.T * g.N * g.Cin, 4.0 * g.B * g.T * (g.N + g.Cin)
        if name in ('tg_gru_layer_fwd', 'tg_gru_layer_fwd_tf32'):
            B, T, H = args[-4], args[-3], args[-2]
            return 2.0 * 2 * B * T * 3 * H * H, 4.0 * B * T * (6 * H + 2 * H * 5)
        if name in ('tg_gru_layer_bwd', 'tg_gru_layer_bwd_tf32'):
            B, T, H = args[-4], args[-3], args[-2]
            return 2.0 * 2 * B * T * 3 * H * H, 4.0 * B * T * (12 * H + 2 * H * 6)
        if name == 'tg_conv1_direct_f32':
            B, Tin, Tout, N, taps = args[4], args[5], args[6], args[7], args[8]
            return 2.0 * B * Tout * N * taps, 4.0 * (B * Tin + B * Tout * N)
        if name == 'tg_adam_flat':
            return 0.0, 28.0 * args[4]
    except Exception:
        pass
    return 0.0, 0.0


class KernelTimer:
    def __init__(self):
        self.records = []          # (name, start_event, end_event, flops, bytes)
        self._real = None

    def __enter__(self):
        real = _lib.load()
        self._real = real
        timer = self

        class Proxy:
            def __getattr__(self, name):
                fn = getattr(real, name)
                if not name.startswith('tg_') or name in ('tg_last_error', 'tg_gru_sync_ints', 'tg_gru_bwd_scratch_floats',
                                                          'tg_device_info', 'tg_version', 'tg_struct_sizes'):
                    return fn

                def timed(*args):
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    rc = fn(*args)
                    e.record()
                    fl, by = _work(name, args)
                    timer.records.append((name, s, e, fl, by, _tag(name, args)))
                    return rc
                return timed
        _lib._lib = Proxy()
        return self

    def __exit__(self, *exc):
        _lib._lib = self._real
        return False

    def summary(self):
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for name, s, e, fl, by, tag in self.records:
            a = agg.setdefault(name, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            a['calls'] += 1
            a['ms'] += s.elapsed_time(e)
            a['flops'] += fl
            a['bytes'] += by
        return agg

    def by_shape(self):
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for name, s, e, fl, by, tag in self.records:
            a = agg.setdefault((name, tag), dict(calls=0, ms=0.0, flops=0.0))
            a['calls'] += 1
            a['ms'] += s.elapsed_time(e)
            a['flops'] += fl
        return agg
