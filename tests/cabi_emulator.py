"""TEST INFRASTRUCTURE ONLY: a NumPy restatement of the *documented semantics* of a few C-ABI entry points (include/tg_b200.h,
csrc/gemm_f32.cu, csrc/elementwise.cu, csrc/autoencoder.cu), operating on host memory through the same raw-pointer arguments.

Purpose: host-side launch plans (which kernel, which strides, which prologue, in which order) are pure Python and can be wrong in ways
the C-ABI kernels' own unit tests cannot see.  With this emulator installed in place of libtg_b200.so, a launch plan runs on CPU
tensors and its RESULT can be compared with the oracle in the `-m "not gpu"` suite - the plan's algebra is checked here, the kernels
themselves are checked on the GPU (tests/test_gpu_*.py).  It is not a fallback: the product never imports tests/, and without this
fixture a CPU tensor raises TgError (tests/test_abi_and_modules.py::test_no_cpu_fallback_without_cuda).

Restated: every entry point the launch plans call, in both arithmetic modes (the tensor-core entries tg_gemm_tf32 / tg_wgrad_tf32 /
tg_gru_layer_*_tf32 are restated by their documented arithmetic in float64, i.e. without the TF32 operand rounding - the emulator checks
which operands, strides and epilogues a plan passes, not the rounding); anything else (debug aids, tg_copy_bytes) raises
NotImplementedError."""
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


def _tf32(a, mode):
    """fp32 -> TF32 operand (10-bit mantissa) as float64: 'trunc' drops the low 13 mantissa bits (what a tensor core fed raw fp32 words
    does), 'rna' rounds to nearest (ties away); None leaves the value alone."""
    if mode is None:
        return a.astype(np.float64)
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    if mode == 'bf16':                        # what-if: bf16 operands (8-bit mantissa), round to nearest
        return ((u + np.uint32(0x8000)) & np.uint32(0xFFFF0000)).view(np.float32).astype(np.float64)
    if mode == 'rna':
        u = u + np.uint32(0x1000)
    return (u & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


class EmuLib:
    """Stands in for the ctypes.CDLL object returned by tgb200._lib.load().  tf32_round: None (default: the tensor-core entries compute
    on the exact fp32 operands) or 'trunc' / 'rna' (their operands are first cut to TF32, to PREDICT the fast mode's error on CPU)."""

    def __init__(self, tf32_round=None):
        self.calls = []
        self.tf32_round = tf32_round

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
        assert 0 < C <= 256 and M > 0, 'TG_REQUIRE of tg_col_stats_f64'
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
        assert 0 < C <= 256, 'TG_REQUIRE of tg_bn_bwd_reduce'
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
        assert all(q % 16 == 0 for q in (p, g, m, v)), 'TG_REQUIRE of tg_adam_flat: 16-byte aligned arenas'
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

    # ---------------------------------------------------------------------------------------- small elementwise (csrc/elementwise.cu)
    def tg_transpose_f32(self, x, out, R, C, stream):
        self.calls.append('tg_transpose_f32')
        _arr(out, R * C)[:] = _arr(x, R * C).reshape(R, C).T.reshape(-1)
        return 0

    def tg_mul(self, a, b, out, n, stream):
        self.calls.append('tg_mul')
        _arr(out, n)[:] = _arr(a, n) * _arr(b, n)
        return 0

    def tg_add(self, a, b, out, n, relu, stream):
        self.calls.append('tg_add')
        v = _arr(a, n) + _arr(b, n)
        _arr(out, n)[:] = np.maximum(v, 0) if relu else v
        return 0

    def tg_relu_mask_bwd(self, dy, y, mask, dx, n, stream):
        self.calls.append('tg_relu_mask_bwd')
        v = np.where(_arr(y, n) > 0, _arr(dy, n), np.float32(0))
        if mask:
            v = v * _arr(mask, n)
        _arr(dx, n)[:] = v
        return 0

    def tg_tcn_res_bwd(self, dxo, xo, x, mask, dpre, dc2, n, stream):
        self.calls.append('tg_tcn_res_bwd')
        XO, X = _arr(xo, n), _arr(x, n)
        dp = np.where(XO > 0, _arr(dxo, n), np.float32(0))
        dc = np.where(XO - X > 0, dp * (_arr(mask, n) if mask else np.float32(1)), np.float32(0))
        _arr(dpre, n)[:] = dp
        _arr(dc2, n)[:] = dc
        return 0

    def tg_sum_halves(self, x, out, M, H, stream):
        self.calls.append('tg_sum_halves')
        X = _arr(x, M * 2 * H).reshape(M, 2 * H)
        _arr(out, M * H)[:] = (X[:, :H] + X[:, H:]).reshape(-1)
        return 0

    def tg_dup_halves(self, d, dx, M, H, stream):
        self.calls.append('tg_dup_halves')
        D = _arr(d, M * H).reshape(M, H)
        _arr(dx, M * 2 * H)[:] = np.concatenate([D, D], axis=1).reshape(-1)
        return 0

    def tg_reparam_fwd(self, mu, lv, eps, z, n, stream):
        self.calls.append('tg_reparam_fwd')
        _arr(z, n)[:] = _arr(mu, n) + _arr(eps, n) * np.exp(np.float32(0.5) * _arr(lv, n))
        return 0

    def tg_reparam_bwd(self, dz, lv, eps, dmu, dlv, n, stream):
        self.calls.append('tg_reparam_bwd')
        _arr(dmu, n)[:] += _arr(dz, n)
        _arr(dlv, n)[:] += _arr(dz, n) * _arr(eps, n) * np.float32(0.5) * np.exp(np.float32(0.5) * _arr(lv, n))
        return 0

    def tg_make_pre_seq(self, target, pre, B, T, D, n_pre, stream):
        self.calls.append('tg_make_pre_seq')
        P = np.zeros((B, T, D + 1), np.float32)
        P[:, :n_pre, :D] = _arr(target, B * T * D).reshape(B, T, D)[:, :n_pre]
        P[:, :n_pre, D] = 1
        _arr(pre, B * T * (D + 1))[:] = P.reshape(-1)
        return 0

    def tg_gru_input_concat(self, pre, audio, text, z, out, B, Ba, T, Dp, Da, Dt, Dz, stream):
        self.calls.append('tg_gru_input_concat')
        bi = np.arange(B) % Ba
        parts = []
        if Dp:
            parts.append(_arr(pre, Ba * T * Dp).reshape(Ba, T, Dp)[bi])
        if Da:
            parts.append(_arr(audio, Ba * T * Da).reshape(Ba, T, Da)[bi])
        if Dt:
            parts.append(_arr(text, B * T * Dt).reshape(B, T, Dt))
        if Dz:
            parts.append(np.repeat(_arr(z, B * Dz).reshape(B, 1, Dz), T, axis=1))
        _arr(out, B * T * (Dp + Da + Dt + Dz))[:] = np.concatenate(parts, axis=2).reshape(-1)
        return 0

    def tg_gru_input_split_bwd(self, din, daudio, dtext, dz, B, T, Dp, Da, Dt, Dz, stream):
        self.calls.append('tg_gru_input_split_bwd')
        D = Dp + Da + Dt + Dz
        X = _arr(din, B * T * D).reshape(B, T, D)
        if Da:
            _arr(daudio, B * T * Da)[:] = X[:, :, Dp:Dp + Da].reshape(-1)
        if Dt:
            _arr(dtext, B * T * Dt)[:] = X[:, :, Dp + Da:Dp + Da + Dt].reshape(-1)
        if Dz:
            _arr(dz, B * Dz)[:] = X[:, :, Dp + Da + Dt:].astype(np.float64).sum(1).astype(np.float32).reshape(-1)
        return 0

    def tg_gather_i64(self, src, idx, out, n, stream):
        self.calls.append('tg_gather_i64')
        ix = _arr(idx, n, ctypes.c_longlong)
        _arr(out, n, ctypes.c_longlong)[:] = _arr(src, int(ix.max()) + 1, ctypes.c_longlong)[ix]
        return 0

    # ---------------------------------------------------------------------------------------- embedding / weight norm
    def tg_embedding_gather(self, table, idx, idx_mod, mask, out, M, E, stream):
        self.calls.append('tg_embedding_gather')
        n_idx = idx_mod if idx_mod > 0 else M
        ix = _arr(idx, n_idx, ctypes.c_longlong)[np.arange(M) % n_idx]
        v = _arr(table, (int(ix.max()) + 1) * E).reshape(-1, E)[ix]
        if mask:
            v = v * _arr(mask, M * E).reshape(M, E)
        _arr(out, M * E)[:] = v.reshape(-1)
        return 0

    def tg_embedding_scatter_add(self, dout, idx, mask, dtable, M, E, stream):
        self.calls.append('tg_embedding_scatter_add')
        ix = _arr(idx, M, ctypes.c_longlong)
        v = _arr(dout, M * E).reshape(M, E)
        if mask:
            v = v * _arr(mask, M * E).reshape(M, E)
        T = _arr(dtable, (int(ix.max()) + 1) * E).reshape(-1, E)
        np.add.at(T, ix, v)
        return 0

    def tg_weight_norm_fwd(self, v, g, w, wT, inv_norm, N, Cin, taps, stream):
        self.calls.append('tg_weight_norm_fwd')
        V = _arr(v, N * Cin * taps).reshape(N, Cin, taps)
        inv = (1.0 / np.sqrt((V.astype(np.float64) ** 2).sum((1, 2)))).astype(np.float32)
        W = V * (_arr(g, N) * inv)[:, None, None]                        # [N, Cin, taps]
        _arr(w, N * Cin * taps)[:] = W.transpose(2, 0, 1).reshape(-1)    # tap-major [taps][N][Cin]
        if wT:
            _arr(wT, N * Cin * taps)[:] = W.transpose(2, 1, 0).reshape(-1)   # [taps][Cin][N]
        _arr(inv_norm, N)[:] = inv
        return 0

    def tg_weight_norm_bwd(self, dw, v, g, inv_norm, dv, dg, N, Cin, taps, stream):
        self.calls.append('tg_weight_norm_bwd')
        V = _arr(v, N * Cin * taps).reshape(N, Cin, taps).astype(np.float64)
        DW = _arr(dw, N * Cin * taps).reshape(taps, N, Cin).transpose(1, 2, 0).astype(np.float64)      # -> [N, Cin, taps]
        inv = _arr(inv_norm, N).astype(np.float64); gg = _arr(g, N).astype(np.float64)
        dot = (DW * V).sum((1, 2))
        _arr(dv, N * Cin * taps)[:] += ((gg * inv)[:, None, None] * (DW - V * (dot * inv * inv)[:, None, None])).astype(np.float32).reshape(-1)
        _arr(dg, N)[:] += (dot * inv).astype(np.float32)
        return 0

    def tg_conv1_direct_f32(self, x, w, bias, y, B, Tin, Tout, N, taps, stride, pad, stream):
        self.calls.append('tg_conv1_direct_f32')
        assert 0 < N <= 32 and taps <= 32, 'TG_REQUIRE of tg_conv1_direct_f32'
        X = _arr(x, B * Tin).reshape(B, Tin).astype(np.float64)
        W = _arr(w, N * taps).reshape(N, taps).astype(np.float64)
        ti = (np.arange(Tout) * stride - pad)[:, None] + np.arange(taps)[None, :]
        ok = (ti >= 0) & (ti < Tin)
        win = np.where(ok[None], X[:, np.where(ok, ti, 0)], 0.0)          # [B, Tout, taps]
        v = win @ W.T
        if bias:
            v = v + _arr(bias, N)[None, None, :]
        _arr(y, B * Tout * N)[:] = v.astype(np.float32).reshape(-1)
        return 0

    # ---------------------------------------------------------------------------------------- GRU recurrence (csrc/gru.cu)
    def tg_gru_sync_ints(self, B, H):
        return 64

    def tg_gru_bwd_scratch_floats(self, B, H):
        return 64

    @staticmethod
    def _sig(v):
        return 1.0 / (1.0 + np.exp(-v))

    def tg_gru_layer_fwd(self, gi, whhT_f, whhT_r, bhh_f, bhh_r, out, saved, qstride, sync, B, T, H, stream):
        self.calls.append('tg_gru_layer_fwd')
        assert 0 < H <= 384 and B > 0 and T > 0, 'TG_REQUIRE of tg_gru_layer_fwd'
        GI = _arr(gi, B * T * 6 * H).reshape(B, T, 6 * H).astype(np.float64)
        OUT = _arr(out, B * T * 2 * H).reshape(B, T, 2 * H)
        S = _arr(saved, 3 * qstride + B * T * 2 * H) if saved else None
        rows = np.arange(B) * T
        for d, (wT, bh) in enumerate(((whhT_f, bhh_f), (whhT_r, bhh_r))):
            rnd = getattr(self, '_gru_round', None)
            WT = _tf32(_arr(wT, H * 3 * H).reshape(H, 3 * H), rnd)
            bhv = _arr(bh, 3 * H).astype(np.float64)
            h = np.zeros((B, H))
            for s in range(T):
                t = s if d == 0 else T - 1 - s
                gh = (_tf32(h.astype(np.float32), rnd) if rnd else h) @ WT + bhv
                g = GI[:, t, d * 3 * H:(d + 1) * 3 * H]
                r = self._sig(g[:, :H] + gh[:, :H]); z = self._sig(g[:, H:2 * H] + gh[:, H:2 * H])
                n = np.tanh(g[:, 2 * H:] + r * gh[:, 2 * H:])
                h = (1 - z) * n + z * h
                OUT[:, t, d * H:(d + 1) * H] = h.astype(np.float32)
                if S is not None:
                    o = ((rows + t) * 2 * H + d * H)[:, None] + np.arange(H)[None, :]
                    for q, val in enumerate((r, z, n, gh[:, 2 * H:])):
                        S[q * qstride + o] = val.astype(np.float32)
        return 0

    def tg_dgru_stack_fwd(self, x, params, masks, outs, saved, qstride, drops, w_out, b_out, w_out2, b_out2, hsum, o1, prob, B, T, I0, H, L,
                          fast, stream):
        """csrc/dgru_stack.cu by its documented semantics: L x [gi = W_ih x + b_ih; recurrence; dropout mask] + sum of the directions +
        Linear(64,1) per frame + Linear(T,1) + sigmoid, parameters read from the flat-arena block."""
        self.calls.append('tg_dgru_stack_fwd')
        assert H == 64 and 1 <= L <= 4 and 1 <= T <= 32 and 1 <= I0 <= 64, 'TG_REQUIRE of tg_dgru_stack_fwd'
        ptrs = lambda a: [None] * 4 if not a else list(ctypes.cast(ctypes.c_void_p(a if isinstance(a, int) else a.value), ctypes.POINTER(ctypes.c_void_p * 4)).contents)
        masks, outs, saved, drops = ptrs(masks), ptrs(outs), ptrs(saved), ptrs(drops)
        G3, rows = 3 * H, np.arange(B) * T
        inp = _arr(x, B * T * I0).reshape(B, T, I0).astype(np.float64)
        off = 0
        for l in range(L):
            K = I0 if l == 0 else 2 * H
            n_all = 2 * G3 * K + 2 * G3 + 2 * G3 * H + 2 * G3
            blk = _arr(params + 4 * off, n_all).astype(np.float64)
            off += n_all
            wih = blk[:2 * G3 * K].reshape(2, G3, K); bih = blk[2 * G3 * K:2 * G3 * K + 2 * G3].reshape(2, G3)
            whh = blk[2 * G3 * K + 2 * G3:2 * G3 * K + 2 * G3 + 2 * G3 * H].reshape(2, G3, H); bhh = blk[-2 * G3:].reshape(2, G3)
            OUT = _arr(outs[l], B * T * 2 * H).reshape(B, T, 2 * H)
            S = _arr(saved[l], 3 * qstride + B * T * 2 * H) if saved[l] else None
            full = np.zeros((B, T, 2 * H))
            for d in range(2):
                gi = inp @ wih[d].T + bih[d]
                h = np.zeros((B, H))
                for s_ in range(T):
                    t = s_ if d == 0 else T - 1 - s_
                    gh = h @ whh[d].T + bhh[d]
                    g = gi[:, t]
                    r = self._sig(g[:, :H] + gh[:, :H]); z = self._sig(g[:, H:2 * H] + gh[:, H:2 * H])
                    n = np.tanh(g[:, 2 * H:] + r * gh[:, 2 * H:])
                    h = (1 - z) * n + z * h
                    full[:, t, d * H:(d + 1) * H] = h
                    if S is not None:
                        o = ((rows + t) * 2 * H + d * H)[:, None] + np.arange(H)[None, :]
                        for q, val in enumerate((r, z, n, gh[:, 2 * H:])):
                            S[q * qstride + o] = val.astype(np.float32)
            OUT[:] = full.astype(np.float32)
            inp = OUT.astype(np.float64)
            if l < L - 1 and masks[l]:
                inp = inp * _arr(masks[l], B * T * 2 * H).reshape(B, T, 2 * H)
                if drops[l]:
                    _arr(drops[l], B * T * 2 * H)[:] = inp.astype(np.float32).reshape(-1)
                    inp = _arr(drops[l], B * T * 2 * H).reshape(B, T, 2 * H).astype(np.float64)
        hs = (inp[:, :, :H] + inp[:, :, H:]).astype(np.float32)
        _arr(hsum, B * T * H)[:] = hs.reshape(-1)
        o = (hs.astype(np.float64) @ _arr(w_out, H).astype(np.float64) + float(_arr(b_out, 1)[0])).astype(np.float32)
        _arr(o1, B * T)[:] = o.reshape(-1)
        v = o.astype(np.float64) @ _arr(w_out2, T).astype(np.float64) + float(_arr(b_out2, 1)[0])
        _arr(prob, B)[:] = (1.0 / (1.0 + np.exp(-v))).astype(np.float32)
        return 0

    def tg_gru_layer_bwd(self, dout, out, saved, qstride, whh_f, whh_r, dgi, dgh, partial, sync, B, T, H, stream):
        self.calls.append('tg_gru_layer_bwd')
        DOUT = _arr(dout, B * T * 2 * H).reshape(B, T, 2 * H).astype(np.float64)
        OUT = _arr(out, B * T * 2 * H).reshape(B, T, 2 * H).astype(np.float64)
        S = _arr(saved, 3 * qstride + B * T * 2 * H)
        DGI = _arr(dgi, B * T * 6 * H).reshape(B, T, 6 * H); DGH = _arr(dgh, B * T * 6 * H).reshape(B, T, 6 * H)
        rows = np.arange(B) * T
        for d, wp in enumerate((whh_f, whh_r)):
            rnd = getattr(self, '_gru_round', None)
            W = _tf32(_arr(wp, 3 * H * H).reshape(3 * H, H), rnd)
            carry = np.zeros((B, H))
            for s in range(T):
                t = T - 1 - s if d == 0 else s
                tp = t - 1 if d == 0 else t + 1
                o = ((rows + t) * 2 * H + d * H)[:, None] + np.arange(H)[None, :]
                r, z, n, hn = (S[q * qstride + o].astype(np.float64) for q in range(4))
                hprev = OUT[:, tp, d * H:(d + 1) * H] if 0 <= tp < T else np.zeros((B, H))
                dh = DOUT[:, t, d * H:(d + 1) * H] + carry
                dn = dh * (1 - z) * (1 - n * n)
                dzp = dh * (hprev - n) * z * (1 - z)
                drp = dn * hn * r * (1 - r)
                dnr = dn * r
                DGI[:, t, d * 3 * H:(d + 1) * 3 * H] = np.concatenate([drp, dzp, dn], axis=1).astype(np.float32)
                gh = np.concatenate([drp, dzp, dnr], axis=1)
                DGH[:, t, d * 3 * H:(d + 1) * 3 * H] = gh.astype(np.float32)
                carry = dh * z + (_tf32(gh.astype(np.float32), rnd) if rnd else gh) @ W
        return 0

    def tg_dgru_stack_bwd(self, dlogit, params, masks, outs, saved, qstride, hsum, o1, w_out, w_out2, dgi, dgh, dx0, g_w_out, g_b_out, g_w_out2,
                          g_b_out2, B, T, I0, H, L, fast, stream):
        """csrc/dgru_stack.cu (backward) by its documented semantics."""
        self.calls.append('tg_dgru_stack_bwd')
        assert H == 64 and 1 <= L <= 4 and 1 <= T <= 32 and 4 <= I0 <= 64, 'TG_REQUIRE of tg_dgru_stack_bwd'
        ptrs = lambda a: [None] * 4 if not a else list(ctypes.cast(ctypes.c_void_p(a if isinstance(a, int) else a.value), ctypes.POINTER(ctypes.c_void_p * 4)).contents)
        masks, outs, saved, dgi, dgh = ptrs(masks), ptrs(outs), ptrs(saved), ptrs(dgi), ptrs(dgh)
        G3, rows = 3 * H, np.arange(B) * T
        dl = _arr(dlogit, B).astype(np.float64)
        O1 = _arr(o1, B * T).reshape(B, T).astype(np.float64)
        HS = _arr(hsum, B * T * H).reshape(B, T, H).astype(np.float64)
        w2 = _arr(w_out2, T).astype(np.float64); wo = _arr(w_out, H).astype(np.float64)
        do1 = dl[:, None] * w2[None, :]
        _arr(g_w_out2, T)[:] += (dl[:, None] * O1).sum(0).astype(np.float32)
        _arr(g_b_out2, 1)[:] += np.float32(dl.sum())
        _arr(g_b_out, 1)[:] += np.float32(do1.sum())
        _arr(g_w_out, H)[:] += np.einsum('bt,bth->h', do1, HS).astype(np.float32)
        dcur = np.concatenate([do1[:, :, None] * wo[None, None, :]] * 2, axis=2)          # [B,T,2H]
        sizes = []
        for l in range(L):
            K = I0 if l == 0 else 2 * H
            sizes.append(2 * G3 * K + 2 * G3 + 2 * G3 * H + 2 * G3)
        for l in range(L - 1, -1, -1):
            K = I0 if l == 0 else 2 * H
            blk = _arr(params + 4 * sum(sizes[:l]), sizes[l]).astype(np.float64)
            wih = blk[:2 * G3 * K].reshape(2 * G3, K)
            whh = blk[2 * G3 * K + 2 * G3:2 * G3 * K + 2 * G3 + 2 * G3 * H].reshape(2, G3, H)
            OUT = _arr(outs[l], B * T * 2 * H).reshape(B, T, 2 * H).astype(np.float64)
            S = _arr(saved[l], 3 * qstride + B * T * 2 * H)
            DGI = _arr(dgi[l], B * T * 6 * H).reshape(B, T, 6 * H); DGH = _arr(dgh[l], B * T * 6 * H).reshape(B, T, 6 * H)
            for d in range(2):
                carry = np.zeros((B, H))
                for s_ in range(T):
                    t = T - 1 - s_ if d == 0 else s_
                    tp = t - 1 if d == 0 else t + 1
                    o = ((rows + t) * 2 * H + d * H)[:, None] + np.arange(H)[None, :]
                    r, z, n, hn = (S[q * qstride + o].astype(np.float64) for q in range(4))
                    hprev = OUT[:, tp, d * H:(d + 1) * H] if 0 <= tp < T else np.zeros((B, H))
                    dh = dcur[:, t, d * H:(d + 1) * H] + carry
                    dn = dh * (1 - z) * (1 - n * n)
                    dzp = dh * (hprev - n) * z * (1 - z)
                    drp = dn * hn * r * (1 - r)
                    dnr = dn * r
                    DGI[:, t, d * 3 * H:(d + 1) * 3 * H] = np.concatenate([drp, dzp, dn], axis=1).astype(np.float32)
                    gh = np.concatenate([drp, dzp, dnr], axis=1)
                    DGH[:, t, d * 3 * H:(d + 1) * 3 * H] = gh.astype(np.float32)
                    carry = dh * z + gh @ whh[d]
            dx = DGI.astype(np.float64) @ wih                                                # [B,T,6H] x [6H,K]
            if l > 0:
                dcur = dx * _arr(masks[l - 1], B * T * 2 * H).reshape(B, T, 2 * H) if masks[l - 1] else dx
            elif dx0:
                _arr(dx0, B * T * I0)[:] = dx.astype(np.float32).reshape(-1)
        return 0

    # ---------------------------------------------------------------------------------------- fused ConvDiscriminator convolutions (csrc/dconv_stack.cu)
    @staticmethod
    def _conv3(a, w, b):
        """channels-last valid convolution, k = 3: a [B,T,Cin], w [N,Cin,3] -> [B,T-2,N]"""
        T = a.shape[1]
        return sum(a[:, j:T - 2 + j, :] @ w[:, :, j].T for j in range(3)) + b

    def tg_dconv_stack_fwd(self, x, w1, b1, g1, be1, rm1, rv1, nbt1, w2, b2, g2, be2, rm2, rv2, nbt2, w3, b3, y0, y1, y2, st1, st2, B, T, D,
                           training, eps, momentum, stream):
        self.calls.append('tg_dconv_stack_fwd')
        assert 0 < B <= 128 and T == 34 and D == 27, 'TG_REQUIRE of tg_dconv_stack_fwd'
        f = lambda p_, n: _arr(p_, n).astype(np.float64)
        a = f(x, B * 34 * 27).reshape(B, 34, 27)
        outs = []
        for (w, b, cin, cout, g, be, rm, rv, nbt, st, yp) in ((w1, b1, 27, 16, g1, be1, rm1, rv1, nbt1, st1, y0), (w2, b2, 16, 8, g2, be2, rm2, rv2, nbt2, st2, y1),
                                                          (w3, b3, 8, 8, None, None, None, None, None, None, y2)):
            y = self._conv3(a, f(w, cout * cin * 3).reshape(cout, cin, 3), f(b, cout))
            _arr(yp, y.size)[:] = y.astype(np.float32).reshape(-1)
            if g is None:
                break
            y = _arr(yp, y.size).astype(np.float64).reshape(y.shape)
            M = y.shape[0] * y.shape[1]
            RM, RV = _arr(rm, cout), _arr(rv, cout)
            if training:
                mu = y.reshape(M, cout).mean(0); var = np.maximum((y.reshape(M, cout) ** 2).mean(0) - mu * mu, 0.0)
                mu32 = mu.astype(np.float32)
                RM[:] = (1.0 - np.float32(momentum)) * RM + np.float32(momentum) * mu32
                RV[:] = (1.0 - np.float32(momentum)) * RV + np.float32(momentum) * (var * M / max(M - 1, 1)).astype(np.float32)
                if nbt:
                    _arr(nbt, 1, ctypes.c_longlong)[0] += 1
                rs = (1.0 / np.sqrt(var + eps)).astype(np.float32)
            else:
                mu32 = RM.copy(); rs = (1.0 / np.sqrt(RV + np.float32(eps))).astype(np.float32)
            sc = _arr(g, cout) * rs
            sh = _arr(be, cout) - mu32 * _arr(g, cout) * rs
            S = _arr(st, 4 * cout)
            S[:cout] = mu32; S[cout:2 * cout] = rs; S[2 * cout:3 * cout] = sc; S[3 * cout:] = sh
            a = y * sc.astype(np.float64) + sh.astype(np.float64)
        return 0

    def tg_dconv_stack_bwd(self, dy2, x, y0, y1, st1, st2, w1, w2, w3, g1, g2, dw1, db1, dw2, db2, dw3, db3, dg1, dbe1, dg2, dbe2, dx, B, T, D, stream):
        self.calls.append('tg_dconv_stack_bwd')
        assert 0 < B <= 128 and T == 34 and D == 27, 'TG_REQUIRE of tg_dconv_stack_bwd'
        f = lambda p_, n: _arr(p_, n).astype(np.float64)
        X = f(x, B * 34 * 27).reshape(B, 34, 27); Y0 = f(y0, B * 32 * 16).reshape(B, 32, 16); Y1 = f(y1, B * 30 * 8).reshape(B, 30, 8)
        S1, S2 = f(st1, 64), f(st2, 32)
        d = f(dy2, B * 28 * 8).reshape(B, 28, 8)

        def conv_bwd(dy, a, w, cin, cout, dwp, dbp, need_da=True):
            W = f(w, cout * cin * 3).reshape(cout, cin, 3)
            Tout = dy.shape[1]
            dW = np.stack([np.einsum('btn,btc->nc', dy, a[:, j:j + Tout, :]) for j in range(3)], axis=2)
            _arr(dwp, dW.size)[:] += dW.astype(np.float32).reshape(-1)
            _arr(dbp, cout)[:] += dy.sum((0, 1)).astype(np.float32)
            if not need_da:
                return None
            da = np.zeros_like(a)
            for j in range(3):
                da[:, j:j + Tout, :] += dy @ W[:, :, j]
            return da

        def bn_bwd(da, y, S, g, C, dgp, dbp):
            mu, rs = S[:C], S[C:2 * C]
            xhat = (y - mu) * rs
            M = y.shape[0] * y.shape[1]
            s0, s1 = da.sum((0, 1)), (da * xhat).sum((0, 1))
            _arr(dgp, C)[:] += s1.astype(np.float32); _arr(dbp, C)[:] += s0.astype(np.float32)
            return f(g, C) * rs * (da - s0 / M - xhat * s1 / M)

        a1 = Y1 * S2[16:24] + S2[24:32]
        da1 = conv_bwd(d, a1, w3, 8, 8, dw3, db3)
        dY1 = bn_bwd(da1, Y1, S2, g2, 8, dg2, dbe2)
        a0 = Y0 * S1[32:48] + S1[48:64]
        da0 = conv_bwd(dY1, a0, w2, 16, 8, dw2, db2)
        dY0 = bn_bwd(da0, Y0, S1, g1, 16, dg1, dbe1)
        dX = conv_bwd(dY0, X, w1, 27, 16, dw1, db1, need_da=bool(dx))
        if dx:
            _arr(dx, B * 34 * 27)[:] = dX.astype(np.float32).reshape(-1)
        return 0

    # ---------------------------------------------------------------------------------------- Speech2Gesture pieces (csrc/s2g.cu)
    @staticmethod
    def _im2col_index(B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo):
        ho, wo, i, j = np.meshgrid(np.arange(Ho), np.arange(Wo), np.arange(kh), np.arange(kw), indexing='ij')
        h = ho * sh + i - pt; w = wo * sw + j - pl
        valid = (h >= 0) & (h < H) & (w >= 0) & (w < W)
        return np.where(valid, h, 0), np.where(valid, w, 0), valid

    def tg_im2col2d(self, x, col, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo, stream):
        self.calls.append('tg_im2col2d')
        X = _arr(x, B * H * W * C).reshape(B, H, W, C)
        h, w, valid = self._im2col_index(B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo)
        out = X[:, h, w, :] * valid[None, ..., None]                       # [B,Ho,Wo,kh,kw,C]
        _arr(col, B * Ho * Wo * kh * kw * C)[:] = out.reshape(-1)
        return 0

    def tg_col2im2d(self, col, dx, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo, stream):
        self.calls.append('tg_col2im2d')
        Cm = _arr(col, B * Ho * Wo * kh * kw * C).reshape(B, Ho, Wo, kh, kw, C).astype(np.float64)
        h, w, valid = self._im2col_index(B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo)
        out = np.zeros((B, H, W, C))
        np.add.at(out, (slice(None), h, w, slice(None)), Cm * valid[None, ..., None])
        _arr(dx, B * H * W * C)[:] = out.astype(np.float32).reshape(-1)
        return 0

    @staticmethod
    def _bilin(o, n_in, n_out):
        s = np.maximum((o + 0.5) * (np.float32(n_in) / np.float32(n_out)) - 0.5, 0.0).astype(np.float32)
        i0 = np.minimum(s.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        return i0, i1, (s - i0).astype(np.float64)

    def _bilin_weights(self, H, W, Ho, Wo):
        h0, h1, lh = self._bilin(np.arange(Ho), H, Ho)
        w0, w1, lw = self._bilin(np.arange(Wo), W, Wo)
        return h0, h1, lh, w0, w1, lw

    def tg_resize_bilinear_fwd(self, x, y, B, H, W, C, Ho, Wo, stream):
        self.calls.append('tg_resize_bilinear_fwd')
        X = _arr(x, B * H * W * C).reshape(B, H, W, C).astype(np.float64)
        h0, h1, lh, w0, w1, lw = self._bilin_weights(H, W, Ho, Wo)
        lh = lh[None, :, None, None]; lw = lw[None, None, :, None]
        g = lambda hh, ww: X[:, hh][:, :, ww]
        out = (1 - lh) * ((1 - lw) * g(h0, w0) + lw * g(h0, w1)) + lh * ((1 - lw) * g(h1, w0) + lw * g(h1, w1))
        _arr(y, B * Ho * Wo * C)[:] = out.astype(np.float32).reshape(-1)
        return 0

    def tg_resize_bilinear_bwd(self, dy, dx, B, H, W, C, Ho, Wo, stream):
        self.calls.append('tg_resize_bilinear_bwd')
        G = _arr(dy, B * Ho * Wo * C).reshape(B, Ho, Wo, C).astype(np.float64)
        h0, h1, lh, w0, w1, lw = self._bilin_weights(H, W, Ho, Wo)
        out = np.zeros((B, H, W, C))
        for hh, wh in ((h0, 1 - lh), (h1, lh)):
            for ww, wl in ((w0, 1 - lw), (w1, lw)):
                np.add.at(out, (slice(None), hh[:, None], ww[None, :], slice(None)), G * wh[None, :, None, None] * wl[None, None, :, None])
        _arr(dx, B * H * W * C)[:] = out.astype(np.float32).reshape(-1)
        return 0

    def tg_upsample2_add_fwd(self, x1, x2, y, B, T1, T2, C, stream):
        self.calls.append('tg_upsample2_add_fwd')
        X1 = _arr(x1, B * T1 * C).reshape(B, T1, C); X2 = _arr(x2, B * T2 * C).reshape(B, T2, C)
        _arr(y, B * T2 * C)[:] = (np.repeat(X1, 2, axis=1)[:, :T2] + X2).reshape(-1)
        return 0

    def tg_upsample2_bwd(self, dy, dx1, B, T1, T2, C, accumulate, stream):
        self.calls.append('tg_upsample2_bwd')
        G = np.zeros((B, 2 * T1, C), np.float32); G[:, :T2] = _arr(dy, B * T2 * C).reshape(B, T2, C)
        v = G.reshape(B, T1, 2, C).sum(2)
        out = _arr(dx1, B * T1 * C)
        out[:] = (out.reshape(B, T1, C) + v if accumulate else v).reshape(-1)
        return 0

    def tg_time_diff_fwd(self, x, y, B, T, D, stream):
        self.calls.append('tg_time_diff_fwd')
        X = _arr(x, B * T * D).reshape(B, T, D)
        _arr(y, B * (T - 1) * D)[:] = (X[:, 1:] - X[:, :-1]).reshape(-1)
        return 0

    def tg_time_diff_bwd(self, dy, dx, B, T, D, accumulate, stream):
        self.calls.append('tg_time_diff_bwd')
        G = _arr(dy, B * (T - 1) * D).reshape(B, T - 1, D)
        v = np.zeros((B, T, D), np.float32); v[:, 1:] += G; v[:, :-1] -= G
        out = _arr(dx, B * T * D)
        out[:] = (out.reshape(B, T, D) + v if accumulate else v).reshape(-1)
        return 0

    def tg_concat_bcast_fwd(self, a, p, y, B, T, Ca, Cp, stream):
        self.calls.append('tg_concat_bcast_fwd')
        A = _arr(a, B * T * Ca).reshape(B, T, Ca); P = _arr(p, B * Cp).reshape(B, 1, Cp)
        _arr(y, B * T * (Ca + Cp))[:] = np.concatenate([A, np.broadcast_to(P, (B, T, Cp))], axis=2).reshape(-1)
        return 0

    def tg_concat_bcast_bwd(self, d, da, dp, B, T, Ca, Cp, stream):
        self.calls.append('tg_concat_bcast_bwd')
        Dd = _arr(d, B * T * (Ca + Cp)).reshape(B, T, Ca + Cp)
        _arr(da, B * T * Ca)[:] = Dd[:, :, :Ca].reshape(-1)
        _arr(dp, B * Cp)[:] = Dd[:, :, Ca:].astype(np.float64).sum(1).astype(np.float32).reshape(-1)
        return 0

    def tg_lrelu_bwd(self, dy, x, dx, n, slope, stream):
        self.calls.append('tg_lrelu_bwd')
        _arr(dx, n)[:] = _arr(dy, n) * np.where(_arr(x, n) >= 0, np.float32(1), np.float32(slope))
        return 0

    def tg_mse_const(self, x, n, target, w, scalar, dx, stream):
        self.calls.append('tg_mse_const')
        e = _arr(x, n).astype(np.float64) - target
        _arr(scalar, 1, ctypes.c_double)[0] += (e * e).mean()
        if dx:
            _arr(dx, n)[:] = (w * 2.0 * e / n).astype(np.float32)
        return 0

    def tg_l1_loss(self, x, y, n, w, scalar, dx, stream):
        self.calls.append('tg_l1_loss')
        e = _arr(x, n).astype(np.float64) - _arr(y, n)
        _arr(scalar, 1, ctypes.c_double)[0] += np.abs(e).mean()
        if dx:
            _arr(dx, n)[:] = (w * np.sign(e) / n).astype(np.float32)
        return 0

    # ---------------------------------------------------------------------------------------- losses (csrc/losses.cu)
    @staticmethod
    def _huber(x, y, beta):
        uu = x / beta - y / beta
        d = np.abs(uu)
        return np.where(d < 1, 0.5 * d * d * beta, (d - 0.5) * beta), np.where(d < 1, uu, np.sign(uu))

    def tg_gen_losses(self, out, target, out_rand, z, z_rand, mu, logvar, B, TD, Z, w_reg, w_div, w_kld, scalars, d_out, dmu, dlogvar, stream):
        self.calls.append('tg_gen_losses')
        o = _arr(out, B * TD).reshape(B, TD).astype(np.float64); tg = _arr(target, B * TD).reshape(B, TD).astype(np.float64)
        sc = _arr(scalars, 3, ctypes.c_double)
        hv, hg = self._huber(o, tg, 0.1)
        sc[0] += hv.sum()
        grad = (w_reg / (B * TD)) * hg
        if out_rand:
            orr = _arr(out_rand, B * TD).reshape(B, TD).astype(np.float64)
            pv, pg = self._huber(o, orr, 0.05)
            zl = np.abs(_arr(z, B * Z).reshape(B, Z).astype(np.float64) - _arr(z_rand, B * Z).reshape(B, Z)).sum(1) / Z
            denom = zl + 1.0e-5
            raw = -(pv.sum(1) / denom)
            sc[1] += np.maximum(raw, -1000.0).sum()
            coef = np.where(raw >= -1000.0, -1.0 / denom, 0.0)
            grad = grad + (w_div * coef / B)[:, None] * pg
        if mu:
            m = _arr(mu, B * Z).astype(np.float64); lv = _arr(logvar, B * Z).astype(np.float64)
            sc[2] += (1 + lv - m * m - np.exp(lv)).sum()
            if dmu:
                _arr(dmu, B * Z)[:] = (w_kld * m / (B * Z)).astype(np.float32)
                _arr(dlogvar, B * Z)[:] = (-0.5 * w_kld * (1 - np.exp(lv)) / (B * Z)).astype(np.float32)
        if d_out:
            _arr(d_out, B * TD)[:] = grad.astype(np.float32).reshape(-1)
        return 0

    def tg_bce_sigmoid(self, p, n, s, o, w, scalar, dlogit, stream):
        self.calls.append('tg_bce_sigmoid')
        pv = _arr(p, n).astype(np.float64)
        a = (s * pv + o) + 1e-8
        _arr(scalar, 1, ctypes.c_double)[0] += (-np.log(a)).sum() / n
        if dlogit:
            _arr(dlogit, n)[:] = (w * (-s / (n * a)) * pv * (1 - pv)).astype(np.float32)
        return 0

    # ---------------------------------------------------------------------------------------- FGD statistics (csrc/elementwise.cu)
    def tg_feature_stats_f64(self, feat, n, F, acc, stream):
        self.calls.append('tg_feature_stats_f64')
        assert 0 < F <= 64 and n > 0, 'TG_REQUIRE of tg_feature_stats_f64'
        X = _arr(feat, n * F).reshape(n, F).astype(np.float64)
        a = _arr(acc, 1 + F + F * F, ctypes.c_double)
        a[0] += n; a[1:1 + F] += X.sum(0); a[1 + F:] += (X.T @ X).reshape(-1)
        return 0

    def tg_l1_dist_f64(self, a, b, n, acc, stream):
        self.calls.append('tg_l1_dist_f64')
        _arr(acc, 1, ctypes.c_double)[0] += np.abs(_arr(a, n) - _arr(b, n)).astype(np.float64).sum()
        return 0

    # ---------------------------------------------------------------------------------------- seq2seq baseline (csrc/seq2seq.cu)
    def tg_gru_gates_fwd(self, gi, ldgi, gh, hprev, lengths, t, hnew, out, ldout, saved, plane, B, H, stream):
        self.calls.append('tg_gru_gates_fwd')
        bi = np.arange(B)[:, None]; j = np.arange(H)[None, :]
        GI = _arr(gi, (B - 1) * ldgi + 3 * H)
        g = [GI[bi * ldgi + q * H + j].astype(np.float64) for q in range(3)]
        GH = _arr(gh, B * 3 * H).reshape(B, 3 * H).astype(np.float64)
        hp = _arr(hprev, B * H).reshape(B, H).astype(np.float64) if hprev else np.zeros((B, H))
        valid = np.ones((B, 1), bool) if not lengths else (t < _arr(lengths, B, ctypes.c_longlong))[:, None]
        r = self._sig(g[0] + GH[:, :H]); z = self._sig(g[1] + GH[:, H:2 * H]); hn = GH[:, 2 * H:]
        n = np.tanh(g[2] + r * hn)
        h = (1 - z) * n + z * hp
        _arr(hnew, B * H)[:] = np.where(valid, h, hp).astype(np.float32).reshape(-1)
        if out:
            _arr(out, (B - 1) * ldout + H)[bi * ldout + j] = np.where(valid, h, 0.0).astype(np.float32)
        if saved:
            S = _arr(saved, 3 * plane + B * H)
            for q, val in enumerate((r, z, n, hn)):
                S[q * plane:q * plane + B * H] = val.astype(np.float32).reshape(-1)
        return 0

    def tg_gru_gates_bwd(self, dh, dadd, ldadd, saved, plane, hprev, lengths, t, dgi, lddgi, dgh, dhprev, B, H, stream):
        self.calls.append('tg_gru_gates_bwd')
        bi = np.arange(B)[:, None]; j = np.arange(H)[None, :]
        valid = np.ones((B, 1), bool) if not lengths else (t < _arr(lengths, B, ctypes.c_longlong))[:, None]
        carry = _arr(dh, B * H).reshape(B, H).astype(np.float64).copy() if dh else np.zeros((B, H))
        g = carry + (_arr(dadd, (B - 1) * ldadd + H)[bi * ldadd + j].astype(np.float64) if dadd else 0.0)
        S = _arr(saved, 3 * plane + B * H)
        r, z, n, hn = (S[q * plane:q * plane + B * H].reshape(B, H).astype(np.float64) for q in range(4))
        hp = _arr(hprev, B * H).reshape(B, H).astype(np.float64) if hprev else np.zeros((B, H))
        dpn = g * (1 - z) * (1 - n * n)
        dpz = g * (hp - n) * z * (1 - z)
        dpr = dpn * hn * r * (1 - r)
        DGI = _arr(dgi, (B - 1) * lddgi + 3 * H)
        for q, val in enumerate((dpr, dpz, dpn)):
            DGI[bi * lddgi + q * H + j] = np.where(valid, val, 0.0).astype(np.float32)
        _arr(dgh, B * 3 * H)[:] = np.where(valid, np.concatenate([dpr, dpz, dpn * r], axis=1), 0.0).astype(np.float32).reshape(-1)
        _arr(dhprev, B * H)[:] = np.where(valid, g * z, carry).astype(np.float32).reshape(-1)
        return 0

    def tg_attn_fwd(self, hq, eproj, enc, v, w, ctx, B, Tm, H, stream):
        self.calls.append('tg_attn_fwd')
        HQ = _arr(hq, B * H).reshape(B, 1, H).astype(np.float64); EP = _arr(eproj, B * Tm * H).reshape(B, Tm, H).astype(np.float64)
        EN = _arr(enc, B * Tm * H).reshape(B, Tm, H).astype(np.float64); V = _arr(v, H).astype(np.float64)
        score = (np.tanh(HQ + EP) * V).sum(2)
        e = np.exp(score - score.max(1, keepdims=True))
        W = e / e.sum(1, keepdims=True)
        _arr(w, B * Tm)[:] = W.astype(np.float32).reshape(-1)
        _arr(ctx, B * H)[:] = (W[:, :, None] * EN).sum(1).astype(np.float32).reshape(-1)
        return 0

    def tg_attn_bwd(self, dctx, w, hq, eproj, enc, v, denc, deproj, dv, dhq, B, Tm, H, stream):
        self.calls.append('tg_attn_bwd')
        DC = _arr(dctx, B * H).reshape(B, 1, H).astype(np.float64); W = _arr(w, B * Tm).reshape(B, Tm).astype(np.float64)
        HQ = _arr(hq, B * H).reshape(B, 1, H).astype(np.float64); EP = _arr(eproj, B * Tm * H).reshape(B, Tm, H).astype(np.float64)
        EN = _arr(enc, B * Tm * H).reshape(B, Tm, H).astype(np.float64); V = _arr(v, H).astype(np.float64)
        dw = (DC * EN).sum(2)
        dscore = W * (dw - (W * dw).sum(1, keepdims=True))
        E = np.tanh(HQ + EP)
        dpre = dscore[:, :, None] * V * (1 - E * E)
        _arr(deproj, B * Tm * H)[:] += dpre.astype(np.float32).reshape(-1)
        _arr(denc, B * Tm * H)[:] += (W[:, :, None] * DC).astype(np.float32).reshape(-1)
        _arr(dhq, B * H)[:] = dpre.sum(1).astype(np.float32).reshape(-1)
        _arr(dv, H)[:] += (dscore[:, :, None] * E).sum((0, 1)).astype(np.float32)
        return 0

    def tg_s2s_loss(self, out, target, loss, dy, B, T, D, w_mse, w_cont, w_var, stream):
        self.calls.append('tg_s2s_loss')
        O = _arr(out, B * T * D).reshape(B, T, D).astype(np.float64); Y = _arr(target, B * T * D).reshape(B, T, D).astype(np.float64)
        inv_n = 1.0 / (B * T * D)
        nrm = np.sqrt((O * O).sum(1))                                            # [B, D]: norm over TIME (train_seq2seq.py:17)
        df = O - Y
        dd = O[:, 1:] - O[:, :-1]
        g = 2 * df * w_mse * inv_n - np.where(nrm[:, None, :] > 0, O / np.where(nrm > 0, nrm, 1.0)[:, None, :], 0.0) * w_var * inv_n
        sg = np.sign(dd) * w_cont * inv_n
        g[:, 1:] += sg
        g[:, :-1] -= sg
        g[:, 0] = 0.0
        _arr(dy, T * B * D)[:] = g.transpose(1, 0, 2).astype(np.float32).reshape(-1)
        _arr(loss, 1, ctypes.c_double)[0] += ((df * df).sum() * w_mse + np.abs(dd).sum() * w_cont - nrm.sum() * w_var) * inv_n
        return 0

    def tg_s2s_gather_inputs(self, poses, outputs, xin, B, T, D, n_pre, stream):
        self.calls.append('tg_s2s_gather_inputs')
        P = _arr(poses, B * T * D).reshape(B, T, D); O = _arr(outputs, B * T * D).reshape(B, T, D)
        X = np.zeros((T, B, D), np.float32)
        for t in range(1, T):
            X[t] = (P if t - 1 < n_pre else O)[:, t - 1]
        _arr(xin, T * B * D)[:] = X.reshape(-1)
        return 0

    def tg_sumsq_f64(self, x, n, out, stream):
        self.calls.append('tg_sumsq_f64')
        _arr(out, 1, ctypes.c_double)[0] += (_arr(x, n).astype(np.float64) ** 2).sum()
        return 0

    def tg_clip_scale(self, x, n, sumsq, max_norm, stream):
        self.calls.append('tg_clip_scale')
        coef = min(1.0, max_norm / (math.sqrt(_arr(sumsq, 1, ctypes.c_double)[0]) + 1e-6))
        if coef < 1.0:
            _arr(x, n)[:] *= np.float32(coef)
        return 0

    def tg_pose_eval_metrics(self, out, target, B, T, D, n_pre, acc, stream):
        self.calls.append('tg_pose_eval_metrics')
        assert D == 27
        d = _arr(out, B * T * D).reshape(B, T, 9, 3).astype(np.float64) - _arr(target, B * T * D).reshape(B, T, 9, 3)
        parent, child = (0, 1, 2, 1, 4, 5, 1, 7, 8), (1, 2, 3, 4, 5, 6, 7, 8, 9)
        blen = (0.26, 0.18, 0.14, 0.22, 0.36, 0.33, 0.22, 0.36, 0.33)
        e = np.zeros((B, T, 10, 3))
        for j in range(9):
            e[:, :, child[j]] = e[:, :, parent[j]] + blen[j] * d[:, :, j]
        a = _arr(acc, 3, ctypes.c_double)
        a[0] += np.abs(_arr(out, B * T * D) - _arr(target, B * T * D)).astype(np.float64).sum()
        a[1] += np.abs(e[:, n_pre:]).sum()
        a[2] += np.abs(e[:, 2:] - 2 * e[:, 1:-1] + e[:, :-2]).sum()
        return 0

    # ---------------------------------------------------------------------------------------- Philox-4x32-10 (csrc/common.cuh)
    @staticmethod
    def _philox(seed, stream_id, n4, offset_ptr):
        """First 4x32-bit block of the counter stream of element group i (subsequence = (stream_id << 40) + i), as uint64 arrays."""
        M = np.uint64(0xFFFFFFFF)
        off = np.uint64(_arr(offset_ptr, 1, ctypes.c_longlong)[0]) if offset_ptr else np.uint64(0)
        sub = (np.uint64(stream_id) << np.uint64(40)) + np.arange(n4, dtype=np.uint64)
        k0, k1 = np.uint64(seed) & M, np.uint64(seed) >> np.uint64(32)
        c0 = np.full(n4, off & M, np.uint64); c1 = np.full(n4, off >> np.uint64(32), np.uint64)
        c2 = sub & M; c3 = sub >> np.uint64(32)
        for _ in range(10):
            p0 = np.uint64(0xD2511F53) * c0; p1 = np.uint64(0xCD9E8D57) * c2
            c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & M, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & M
            k0 = (k0 + np.uint64(0x9E3779B9)) & M; k1 = (k1 + np.uint64(0xBB67AE85)) & M
        return c0, c1, c2, c3

    @staticmethod
    def _unit(x):
        return ((x >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0)).astype(np.float32)

    def tg_philox_normal(self, out, n, seed, offset_dev, stream_id, stream):
        self.calls.append('tg_philox_normal')
        u = [self._unit(c) for c in self._philox(seed, stream_id, (n + 3) >> 2, offset_dev)]
        two_pi = np.float32(6.2831853071795864)
        r0 = np.sqrt(np.float32(-2) * np.log(np.float32(1) - u[0])); r1 = np.sqrt(np.float32(-2) * np.log(np.float32(1) - u[2]))
        o = np.stack([r0 * np.cos(two_pi * u[1]), r0 * np.sin(two_pi * u[1]), r1 * np.cos(two_pi * u[3]), r1 * np.sin(two_pi * u[3])], axis=1)
        _arr(out, n)[:] = o.reshape(-1)[:n].astype(np.float32)
        return 0

    def tg_philox_dropout_mask(self, out, n, p, seed, offset_dev, stream_id, stream):
        self.calls.append('tg_philox_dropout_mask')
        u = np.stack([self._unit(c) for c in self._philox(seed, stream_id, (n + 3) >> 2, offset_dev)], axis=1).reshape(-1)[:n]
        pf = np.float32(p)
        _arr(out, n)[:] = np.where(u >= pf, np.float32(1) / (np.float32(1) - pf), np.float32(0))
        return 0

    def tg_philox_randperm(self, out, n, seed, offset_dev, stream_id, stream):
        self.calls.append('tg_philox_randperm')
        assert 0 < n <= 2048, 'TG_REQUIRE of tg_philox_randperm'
        c0, c1, _, _ = self._philox(seed, stream_id, n, offset_dev)
        keys = (c0 << np.uint64(32)) | (c1 & np.uint64(0xFFFFF000)) | np.arange(n, dtype=np.uint64)
        _arr(out, n, ctypes.c_longlong)[:] = (np.sort(keys) & np.uint64(0xFFF)).astype(np.int64)
        return 0

    # ---------------------------------------------------------------------------------------- tensor-core entries (documented arithmetic)
    def _epilogue(self, p, v, rows, N):
        n = np.arange(N)
        if p.escale:
            v = v * _arr(p.escale, N)[None, :]
        if p.bias:
            v = v + _arr(p.bias, N)[None, :]
        v = _act(v, p.act1, p.slope1)
        if p.mask:
            mo = rows[:, None] * p.ldmask + n[None, :]
            v = v * _arr(p.mask, mo.max() + 1)[mo]
        if p.residual:
            ro = rows[:, None] * p.ldres + n[None, :]
            v = v + _arr(p.residual, ro.max() + 1)[ro]
        return _act(v, p.act2, 0.0)

    def tg_gemm_tf32(self, pref, stream):
        p = pref._obj
        self.calls.append('tg_gemm_tf32')
        # TMA descriptors: 16-byte aligned bases and pitches; epilogue activations none / relu / leaky-relu(0..1); clip-mode contract
        assert p.A % 16 == 0 and p.Bw % 16 == 0 and p.lda % 4 == 0 and p.ldb % 4 == 0, 'tg_gemm_tf32: TMA alignment'
        assert p.taps in (1, 2) and 0 <= p.act1 <= 2 and p.act2 in (0, 1) and (p.act1 != 2 or 0.0 <= p.slope1 <= 1.0)
        assert p.clip_rows == 0 or (p.taps == 1 and p.M % p.clip_rows == 0 and p.a_clip_pitch > 0)
        M, N, K = p.M, p.N, p.K
        m = np.arange(M); k = np.arange(K)
        Bw = _arr(p.Bw, (p.taps * N - 1) * p.ldb + K)
        acc = np.zeros((M, N))
        for tap in range(p.taps):
            shift = p.shift0 if (p.taps == 2 and tap == 0) else 0
            if p.clip_rows > 0:
                assert p.taps == 1
                off = ((m // p.clip_rows) * p.a_clip_pitch + (m % p.clip_rows) * p.lda)[:, None] + k[None, :]
                ok = np.ones(M, bool)
            else:
                rows = m + shift
                ok = (rows >= 0) & (rows < p.a_rows)
                if shift != 0:
                    t = m % p.T
                    ok &= (t + shift >= 0) & (t + shift < p.T)
                off = np.where(ok, rows, 0)[:, None] * p.lda + k[None, :]
            A = np.where(ok[:, None], _tf32(_arr(p.A, off.max() + 1)[off], self.tf32_round), 0.0)
            wo = (tap * N + np.arange(N))[:, None] * p.ldb + k[None, :]
            acc += A @ _tf32(Bw[wo], self.tf32_round).T
        v = self._epilogue(p, acc, m, N)
        yo = m[:, None] * p.ldc + np.arange(N)[None, :]
        Y = _arr(p.C, yo.max() + 1)
        if p.accumulate:
            v = v + Y[yo]
        Y[yo] = v.astype(np.float32)
        return 0

    def tg_col_sum_f32(self, g, ld, M, N, out, stream):
        self.calls.append('tg_col_sum_f32')
        o = np.arange(M)[:, None] * ld + np.arange(N)[None, :]
        _arr(out, N)[:] += _arr(g, o.max() + 1)[o].astype(np.float64).sum(0).astype(np.float32)
        return 0

    def tg_wgrad_tf32(self, pref, stream):
        p = pref._obj
        self.calls.append('tg_wgrad_tf32')
        assert p.G % 16 == 0 and p.X % 16 == 0 and p.ldg % 4 == 0 and p.ldx % 4 == 0, 'tg_wgrad_tf32: TMA alignment'
        assert p.shift == 0 or p.B == 1 or p.x_clip_pitch > 0 or p.T <= 40, 'TG_REQUIRE of tg_wgrad_tf32: shifted taps need T <= 40'
        B, T, N, C = p.B, p.T, p.N, p.Cin
        if p.dbias:
            self.tg_col_sum_f32(p.G, p.ldg, B * T, N, p.dbias, stream)
        b = np.repeat(np.arange(B), T); t = np.tile(np.arange(T), B)
        go = (b * T + t)[:, None] * p.ldg + np.arange(N)[None, :]
        G = _tf32(_arr(p.G, go.max() + 1)[go], self.tf32_round)
        ts = t + p.shift
        ok = (ts >= 0) & (ts < T)
        base = b * p.x_clip_pitch + np.where(ok, ts, 0) * p.ldx if p.x_clip_pitch > 0 else (b * T + np.where(ok, ts, 0)) * p.ldx
        xo = base[:, None] + np.arange(C)[None, :]
        X = np.where(ok[:, None], _tf32(_arr(p.X, xo.max() + 1)[xo], self.tf32_round), 0.0)
        wo = np.arange(N)[:, None] * p.ldw + np.arange(C)[None, :]
        dW = _arr(p.dW, wo.max() + 1)
        dW[wo] = (dW[wo] + G.T @ X).astype(np.float32)
        return 0

    def tg_gru_tf32_sync_ints(self, B, H):
        return 64

    def tg_gru_bwd_tf32_scratch_floats(self, B, H):
        return 64

    def tg_gru_layer_fwd_tf32(self, gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, qstride, sync, B, T, H, stream):
        """Weight layout swapped w.r.t. the fp32 entry: weight_hh as stored [3H,H]."""
        assert 32 <= H <= 384 and H % 4 == 0 and whh_f % 16 == 0 and whh_r % 16 == 0 and out % 16 == 0, 'tg_gru_layer_fwd_tf32: H range / TMA alignment'
        tr = [np.ascontiguousarray(_arr(w, 3 * H * H).reshape(3 * H, H).T) for w in (whh_f, whh_r)]
        self._gru_round = self.tf32_round
        try:
            rc = self.tg_gru_layer_fwd(gi, tr[0].ctypes.data, tr[1].ctypes.data, bhh_f, bhh_r, out, saved, qstride, sync, B, T, H, stream)
        finally:
            self._gru_round = None
        self.calls[-1] = 'tg_gru_layer_fwd_tf32'
        return rc

    def tg_gru_layer_fwd_tf32_drop(self, gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, qstride, mask, drop, sync, B, T, H, stream):
        """tg_gru_layer_fwd_tf32 + drop = out * mask"""
        assert mask and drop and mask % 16 == 0 and drop % 16 == 0, 'tg_gru_layer_fwd_tf32_drop: mask / drop'
        rc = self.tg_gru_layer_fwd_tf32(gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, qstride, sync, B, T, H, stream)
        self.calls[-1] = 'tg_gru_layer_fwd_tf32_drop'
        n = B * T * 2 * H
        _arr(drop, n)[:] = _arr(out, n) * _arr(mask, n)
        return rc

    def tg_gru_layer_bwd_tf32(self, dout, out, saved, qstride, whhT_f, whhT_r, dgi, dgh, partial, sync, B, T, H, stream):
        """Takes the transposed recurrent weights [H,3H]."""
        assert 32 <= H <= 384 and H % 4 == 0 and whhT_f % 16 == 0 and whhT_r % 16 == 0, 'tg_gru_layer_bwd_tf32: H range / TMA alignment'
        tr = [np.ascontiguousarray(_arr(w, 3 * H * H).reshape(H, 3 * H).T) for w in (whhT_f, whhT_r)]
        self._gru_round = self.tf32_round
        try:
            rc = self.tg_gru_layer_bwd(dout, out, saved, qstride, tr[0].ctypes.data, tr[1].ctypes.data, dgi, dgh, partial, sync, B, T, H, stream)
        finally:
            self._gru_round = None
        self.calls[-1] = 'tg_gru_layer_bwd_tf32'
        return rc

    # ---------------------------------------------------------------------------------------- WavEncoder fast path (csrc/wav_fast.cu)
    def tg_window_weights(self, w, w2, w2t, N, Cin, k, stream):
        self.calls.append('tg_window_weights')
        W2 = _arr(w, N * Cin * k).reshape(N, Cin, k).transpose(0, 2, 1).reshape(N, k * Cin)
        _arr(w2, N * Cin * k)[:] = W2.reshape(-1)
        if w2t:
            _arr(w2t, N * Cin * k)[:] = W2.T.reshape(-1)
        return 0

    def tg_window_wgrad_add(self, dw2, dw, N, Cin, k, stream):
        self.calls.append('tg_window_wgrad_add')
        _arr(dw, N * Cin * k)[:] += _arr(dw2, N * Cin * k).reshape(N, k, Cin).transpose(0, 2, 1).reshape(-1)
        return 0

    def tg_window_dgrad_weights(self, w, wd, N, Cin, k, stride, stream):
        self.calls.append('tg_window_dgrad_weights')
        ntap = -(-k // stride)
        W = _arr(w, N * Cin * k).reshape(N, Cin, k)
        WD = np.zeros((ntap, stride, Cin, N), np.float32)
        for j in range(ntap):
            for r in range(stride):
                if r + stride * j < k:
                    WD[j, r] = W[:, :, r + stride * j].T
        _arr(wd, ntap * stride * Cin * N)[:] = WD.reshape(-1)
        return 0

    def tg_conv_dgrad_tf32(self, dy, wd, da, B, Tin, Tout, Cin, N, k, stride, stream):
        self.calls.append('tg_conv_dgrad_tf32')
        assert Cin % 4 == 0 and N % 4 == 0 and N >= 8 and Tin >= (Tout - 1) * stride + k, 'TG_REQUIRE of tg_conv_dgrad_tf32'
        assert dy % 16 == 0 and wd % 16 == 0 and da % 16 == 0, 'TG_REQUIRE of tg_conv_dgrad_tf32(alignment)'
        ntap = -(-k // stride)
        Q = -(-Tin // stride)
        DY = _tf32(_arr(dy, B * Tout * N).reshape(B, Tout, N), self.tf32_round).astype(np.float64)
        WD = _tf32(_arr(wd, ntap * stride * Cin * N).reshape(ntap, stride * Cin, N), self.tf32_round).astype(np.float64)
        out = np.zeros((B, Q, stride * Cin))
        for j in range(ntap):
            sh = np.zeros((B, Q, N))                       # dy rows q - j, zero outside [0, Tout)
            lo, hi = j, min(Q, Tout + j)
            if hi > lo:
                sh[:, lo:hi] = DY[:, lo - j:hi - j]
            out += sh @ WD[j].T
        _arr(da, B * Tin * Cin)[:] = out.reshape(B, Q * stride, Cin)[:, :Tin].astype(np.float32).reshape(-1)
        return 0

    def tg_col2im(self, col, da, B, Tin, Tout, Cin, k, stride, stream):
        self.calls.append('tg_col2im')
        assert Cin % 4 == 0 and col % 16 == 0 and da % 16 == 0, 'TG_REQUIRE of tg_col2im'
        COL = _arr(col, B * Tout * k * Cin).reshape(B, Tout, k, Cin).astype(np.float64)
        DA = np.zeros((B, Tin, Cin))
        for t in range(Tout):
            for j in range(k):
                s_ = t * stride + j
                if s_ < Tin:
                    DA[:, s_] += COL[:, t, j]
        _arr(da, B * Tin * Cin)[:] = DA.astype(np.float32).reshape(-1)
        return 0

    def tg_conv1_wgrad(self, x, dy, dW, dbias, B, Tin, Tout, N, taps, stride, pad, stream):
        self.calls.append('tg_conv1_wgrad')
        assert N == 16 and 1 <= taps <= 15 and 1 <= stride <= 8 and dy % 16 == 0, 'TG_REQUIRE of tg_conv1_wgrad'
        X = _arr(x, B * Tin).reshape(B, Tin).astype(np.float64)
        DY = _arr(dy, B * Tout * N).reshape(B, Tout, N).astype(np.float64)
        ti = (np.arange(Tout) * stride - pad)[:, None] + np.arange(taps)[None, :]
        ok = (ti >= 0) & (ti < Tin)
        win = np.where(ok[None], X[:, np.where(ok, ti, 0)], 0.0)          # [B, Tout, taps]
        _arr(dW, N * taps)[:] += np.einsum('btn,btj->nj', DY, win).astype(np.float32).reshape(-1)
        if dbias:
            _arr(dbias, N)[:] += DY.sum((0, 1)).astype(np.float32)
        return 0


class installed:
    """Context manager: routes tgb200's C-ABI calls to an EmuLib and lets CPU tensors through (trace-mode plumbing)."""

    def __init__(self, tf32_round=None):
        self.tf32_round = tf32_round

    def __enter__(self):
        from tgb200 import _lib
        self._lib = _lib
        self._saved = (_lib._lib, _lib.TRACE_ONLY)
        self.emu = EmuLib(self.tf32_round)
        _lib._lib, _lib.TRACE_ONLY = self.emu, True
        return self.emu

    def __exit__(self, *exc):
        self._lib._lib, self._lib.TRACE_ONLY = self._saved
        return False
