"""TEST INFRASTRUCTURE ONLY: a NumPy restatement of the *documented semantics* of a few C-ABI entry points (include/tg_b200.h,
csrc/gemm_f32.cu, csrc/elementwise.cu, csrc/autoencoder.cu), operating on host memory through the same raw-pointer arguments.

Purpose: host-side launch plans (which kernel, which strides, which prologue, in which order) are pure Python and can be wrong in ways
the C-ABI kernels' own unit tests cannot see.  With this emulator installed in place of libtg_b200.so, a launch plan runs on CPU
tensors and its RESULT can be compared with the oracle in the `-m "not gpu"` suite - the plan's algebra is checked here, the kernels
themselves are checked on the GPU (tests/test_gpu_*.py).  It is not a fallback: the product never imports tests/, and without this
fixture a CPU tensor raises TgError (tests/test_abi_and_modules.py::test_no_cpu_fallback_without_cuda).

Only the entry points the auto-encoder trainer and the eval-mode EmbeddingNet forward use are restated; anything else raises."""
import ctypes
import math

import numpy as np


def _arr(ptr, n, ctype=ctypes.c_float):
    return np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(ptr), ctypes.POINTER(ctype)), shape=(int(n),))


def _lrelu(v, slope):
    return np.where(v >= 0, v, v * slope)


def _act(v, act, slope):
    if act == 1:
        return np.maximum(v, 0)
    if act == 2:
        return _lrelu(v, slope)
    if act == 3:
        return 1.0 / (1.0 + np.exp(-v))
    return v


class EmuLib:
    """Stands in for the ctypes.CDLL object returned by tgb200._lib.load()."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name.startswith('tg_'):
            raise NotImplementedError('cabi_emulator: %s is not restated' % name)
        raise AttributeError(name)

    def tg_last_error(self):
        return b'(emulator)'

    # ---------------------------------------------------------------------------------------- implicit GEMM (csrc/gemm_f32.cu)
    def tg_conv_gemm_f32(self, pref, stream):
        p = pref._obj
        self.calls.append('tg_conv_gemm_f32')
        M, K, N = p.B * p.Tout, p.taps * p.Cin, p.N
        m = np.arange(M); b, t = m // p.Tout, m % p.Tout
        kg = np.arange(K); j, c = kg // p.Cin, kg % p.Cin
        tin = (t * p.stride - p.pad)[:, None] + (j * p.dil)[None, :]
        valid = (tin >= 0) & (tin < p.Tin)
        abst = p.a_bstride if p.a_bstride else p.Tin * p.lda
        asc = p.asc if p.asc else 1
        off = np.where(valid, b[:, None] * abst + tin * p.lda + c[None, :] * asc, 0)
        A = _arr(p.A, off.max() + 1)[off].astype(np.float64)
        if p.pscale:
            A = _lrelu(A * _arr(p.pscale, p.Cin)[c][None, :].astype(np.float64) + _arr(p.pshift, p.Cin)[c][None, :], p.pslope)
        A = np.where(valid, A, 0.0)
        n = np.arange(N)
        woff = n[:, None] * p.ldw + (j * p.wsj + c * p.wsc)[None, :]
        W = _arr(p.W, woff.max() + 1)[woff].astype(np.float64)
        v = A @ W.T
        if p.escale:
            v = v * _arr(p.escale, N)[None, :]
        if p.bias:
            v = v + _arr(p.bias, N)[None, :]
        v = _act(v, p.act1, p.slope1)
        orow = b * p.ToutFull + t * p.ostride + p.ooff
        if p.mask:
            mo = orow[:, None] * p.ldmask + n[None, :]
            v = v * _arr(p.mask, mo.max() + 1)[mo]
        if p.residual:
            ro = orow[:, None] * p.ldres + n[None, :]
            v = v + _arr(p.residual, ro.max() + 1)[ro]
        v = _act(v, p.act2, 0.0)
        yo = orow[:, None] * p.ldc + n[None, :]
        Y = _arr(p.Y, yo.max() + 1)
        if p.accumulate:
            v = v + Y[yo]
        Y[yo] = v.astype(np.float32)
        return 0

    def tg_conv_wgrad_f32(self, pref, stream):
        p = pref._obj
        self.calls.append('tg_conv_wgrad_f32')
        M, KW, N = p.B * p.Tout, p.taps * p.Cin, p.N
        m = np.arange(M); b, t = m // p.Tout, m % p.Tout
        kw = np.arange(KW); j, c = kw // p.Cin, kw % p.Cin
        tin = (t * p.stride - p.pad)[:, None] + (j * p.dil)[None, :]
        valid = (tin >= 0) & (tin < p.Tin)
        off = np.where(valid, (b[:, None] * p.Tin + tin) * p.lda + c[None, :], 0)
        A = _arr(p.A, off.max() + 1)[off].astype(np.float64)
        if p.pscale:
            A = _lrelu(A * _arr(p.pscale, p.Cin)[c][None, :].astype(np.float64) + _arr(p.pshift, p.Cin)[c][None, :], p.pslope)
        A = np.where(valid, A, 0.0)
        n = np.arange(N)
        go = m[:, None] * p.ldg + n[None, :]
        G = _arr(p.G, go.max() + 1)[go].astype(np.float64)
        d = G.T @ A                                                     # [N, KW]
        wo = n[:, None] * p.ldw + (j * p.wsj + c * p.wsc)[None, :]
        dW = _arr(p.dW, wo.max() + 1)
        assert len(np.unique(wo)) == wo.size, 'weight-gradient strides alias'
        dW[wo] = (dW[wo] + d).astype(np.float32)
        if p.dbias:
            db = _arr(p.dbias, N)
            db[:] = (db + G.sum(0)).astype(np.float32)
        return 0

    # ---------------------------------------------------------------------------------------- BatchNorm (csrc/elementwise.cu)
    def tg_col_stats_f64(self, x, ld, M, C, sums, stream):
        self.calls.append('tg_col_stats_f64')
        X = _arr(x, (M - 1) * ld + C)
        o = np.arange(M)[:, None] * ld + np.arange(C)[None, :]
        v = X[o].astype(np.float64)
        s = _arr(sums, 2 * C, ctypes.c_double)
        s[:C] += v.sum(0); s[C:] += (v * v).sum(0)
        return 0

    def tg_bn_finalize(self, sums, M, C, eps, momentum, n_updates, gamma, beta, rm, rv, nbt, mean, rstd, scale, shift, stream):
        self.calls.append('tg_bn_finalize')
        s = _arr(sums, 2 * C, ctypes.c_double)
        mu = s[:C] / M
        var = np.maximum(s[C:] / M - mu * mu, 0.0)
        rs = (1.0 / np.sqrt(var + np.float64(np.float32(eps)))).astype(np.float32)
        g = _arr(gamma, C) if gamma else np.ones(C, np.float32)
        bt = _arr(beta, C) if beta else np.zeros(C, np.float32)
        _arr(mean, C)[:] = mu.astype(np.float32); _arr(rstd, C)[:] = rs
        _arr(scale, C)[:] = g * rs
        _arr(shift, C)[:] = bt - mu.astype(np.float32) * g * rs
        if nbt:
            _arr(nbt, 1, ctypes.c_longlong)[0] += n_updates
        if rm:
            unb = (var * (M / max(M - 1, 1))).astype(np.float32)
            RM, RV = _arr(rm, C), _arr(rv, C)
            mom = np.float32(momentum)
            for _ in range(n_updates):
                RM[:] = (np.float32(1) - mom) * RM + mom * mu.astype(np.float32)
                RV[:] = (np.float32(1) - mom) * RV + mom * unb
        return 0

    def tg_bn_eval_fold(self, gamma, beta, rm, rv, eps, conv_bias, scale, shift, C, stream):
        self.calls.append('tg_bn_eval_fold')
        sc = _arr(gamma, C) / np.sqrt(_arr(rv, C) + np.float32(eps))
        cb = _arr(conv_bias, C) if conv_bias else 0.0
        _arr(scale, C)[:] = sc
        _arr(shift, C)[:] = _arr(beta, C) + (cb - _arr(rm, C)) * sc
        return 0

    def tg_affine_lrelu(self, x, y, M, C, scale, shift, slope, stream):
        self.calls.append('tg_affine_lrelu')
        v = _arr(x, M * C).reshape(M, C) * _arr(scale, C)[None, :] + _arr(shift, C)[None, :]
        _arr(y, M * C)[:] = _lrelu(v, np.float32(slope)).reshape(-1)
        return 0

    def _dz(self, dy, x, M, C, scale, shift, slope):
        X = _arr(x, M * C).reshape(M, C).astype(np.float64)
        z = X * _arr(scale, C)[None, :] + _arr(shift, C)[None, :]
        return X, _arr(dy, M * C).reshape(M, C).astype(np.float64) * np.where(z >= 0, 1.0, slope)

    def tg_bn_bwd_reduce(self, dy, x, M, C, mean, rstd, scale, shift, slope, sums, stream):
        self.calls.append('tg_bn_bwd_reduce')
        X, dz = self._dz(dy, x, M, C, scale, shift, slope)
        s = _arr(sums, 2 * C, ctypes.c_double)
        s[:C] += dz.sum(0)
        s[C:] += (dz * (X - _arr(mean, C)[None, :]) * _arr(rstd, C)[None, :]).sum(0)
        return 0

    def tg_bn_bwd_apply(self, dy, x, dx, M, C, mean, rstd, scale, shift, slope, gamma, sums, dgamma, dbeta, stream):
        self.calls.append('tg_bn_bwd_apply')
        X, dz = self._dz(dy, x, M, C, scale, shift, slope)
        s = _arr(sums, 2 * C, ctypes.c_double)
        rs = _arr(rstd, C).astype(np.float64)[None, :]
        xh = (X - _arr(mean, C)[None, :]) * rs
        g = _arr(gamma, C)[None, :] if gamma else 1.0
        _arr(dx, M * C)[:] = (g * rs * (dz - s[:C][None, :] / M - xh * s[C:][None, :] / M)).astype(np.float32).reshape(-1)
        if dgamma:
            _arr(dgamma, C)[:] += s[C:].astype(np.float32)
        if dbeta:
            _arr(dbeta, C)[:] += s[:C].astype(np.float32)
        return 0

    # ---------------------------------------------------------------------------------------- optimiser (csrc/elementwise.cu)
    def tg_increment_i64(self, x, by, stream):
        self.calls.append('tg_increment_i64')
        _arr(x, 1, ctypes.c_longlong)[0] += by
        return 0

    def tg_adam_flat(self, p, g, m, v, n, lr, b1, b2, eps, gscale, step_dev, stream):
        self.calls.append('tg_adam_flat')
        step = float(_arr(step_dev, 1, ctypes.c_longlong)[0])
        f = np.float32
        bc1 = f(1.0 - math.pow(float(f(b1)), step)); bc2s = f(math.sqrt(1.0 - math.pow(float(f(b2)), step)))
        P, G, Mm, V = _arr(p, n), _arr(g, n), _arr(m, n), _arr(v, n)
        gr = G * f(gscale)
        Mm[:] = f(b1) * Mm + (f(1) - f(b1)) * gr
        V[:] = f(b2) * V + (f(1) - f(b2)) * gr * gr
        P[:] = P - (f(lr) / bc1) * Mm / (np.sqrt(V) / bc2s + f(eps))
        return 0

    # ---------------------------------------------------------------------------------------- csrc/autoencoder.cu
    def tg_ae_recon_loss(self, recon, target, B, T, D, use_diff, weight, acc, d_recon, stream):
        self.calls.append('tg_ae_recon_loss')
        r = _arr(recon, B * T * D).reshape(B, T, D); y = _arr(target, B * T * D).reshape(B, T, D)
        u = r - y
        l0 = np.abs(u).astype(np.float64).sum((1, 2)) / (T * D)
        g = np.sign(u) / np.float32(T * D)
        l1 = np.zeros(B)
        if use_diff and T > 1:
            e = (r[:, 1:] - r[:, :-1]) - (y[:, 1:] - y[:, :-1])
            l1 = np.abs(e).astype(np.float64).sum((1, 2)) / ((T - 1) * D)
            s = np.sign(e) / np.float32((T - 1) * D)
            g[:, 1:] += s
            g[:, :-1] -= s
        a = _arr(acc, 2, ctypes.c_double)
        a[0] += (l0 + l1).sum(); a[1] += l0.sum()
        if d_recon:
            _arr(d_recon, B * T * D)[:] = (np.float32(weight) * g).astype(np.float32).reshape(-1)
        return 0

    def tg_transpose_batched_f32(self, x, out, B, R, C, stream):
        self.calls.append('tg_transpose_batched_f32')
        _arr(out, B * R * C)[:] = _arr(x, B * R * C).reshape(B, R, C).transpose(0, 2, 1).reshape(-1)
        return 0


class installed:
    """Context manager: routes tgb200's C-ABI calls to an EmuLib and lets CPU tensors through (trace-mode plumbing)."""

    def __enter__(self):
        from tgb200 import _lib
        self._lib = _lib
        self._saved = (_lib._lib, _lib.TRACE_ONLY)
        self.emu = EmuLib()
        _lib._lib, _lib.TRACE_ONLY = self.emu, True
        return self.emu

    def __exit__(self, *exc):
        self._lib._lib, self._lib.TRACE_ONLY = self._saved
        return False
