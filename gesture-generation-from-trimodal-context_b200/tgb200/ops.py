"""Tensor-level wrappers over the C ABI (include/tg_b200.h).  Every function enqueues hand-written sm_100a kernels on
torch's current CUDA stream; torch is used only for memory (tensors) and streams.  CUDA tensors only - no fallback."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import ConvGemm, ConvWgrad, GemmTf32, WgradTf32, check

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3

_launches = 0           # kernel launches issued through this module (bench.py reports it as gpu_launches)


def launches() -> int:
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _L():
    return _lib.load()


def _s():
    if _lib.TRACE_ONLY:
        return 0
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('tgb200 ops take CUDA tensors only (no CPU fallback)')
    return t.data_ptr()


def _f32(t: torch.Tensor):
    assert t.dtype == torch.float32, t.dtype
    return t


def device_info():
    out = (ctypes.c_int * 2)()
    _L().tg_device_info(out)
    return out[0], out[1]


# ------------------------------------------------------------------------------------------ implicit GEMM
def conv_gemm(A, W, Y, *, B, Tin, Tout, N, Cin, taps=1, stride=1, dil=1, pad=0, lda=None, ldw=None, wsj=0, wsc=1, ldc=None,
              ToutFull=None, ostride=1, ooff=0, pscale=None, pshift=None, pslope=1.0, escale=None, bias=None, act1=0,
              slope1=0.0, mask=None, ldmask=None, residual=None, ldres=None, act2=0, accumulate=False, asc=1, a_bstride=0,
              w_off=0):
    g = ConvGemm()
    g.A = _p(_f32(A)); g.lda = Cin if lda is None else lda; g.asc = asc; g.a_bstride = a_bstride
    g.W = _p(_f32(W)) + 4 * w_off; g.ldw = taps * Cin if ldw is None else ldw; g.wsj = wsj; g.wsc = wsc
    g.Y = _p(_f32(Y)); g.ldc = N if ldc is None else ldc
    g.B, g.Tout, g.Tin, g.N, g.Cin, g.taps, g.stride, g.dil, g.pad = B, Tout, Tin, N, Cin, taps, stride, dil, pad
    g.ToutFull = Tout if ToutFull is None else ToutFull; g.ostride = ostride; g.ooff = ooff
    g.pscale = _p(pscale); g.pshift = _p(pshift); g.pslope = pslope
    g.escale = _p(escale); g.bias = _p(bias); g.act1 = act1; g.slope1 = slope1
    g.mask = _p(mask); g.ldmask = (N if ldmask is None else ldmask)
    g.residual = _p(residual); g.ldres = (N if ldres is None else ldres)
    g.act2 = act2; g.accumulate = 1 if accumulate else 0
    check(_L().tg_conv_gemm_f32(ctypes.byref(g), _s()), 'tg_conv_gemm_f32')
    _count()


def conv_wgrad(A, G, dW, *, B, Tin, Tout, N, Cin, taps=1, stride=1, dil=1, pad=0, lda=None, ldg=None, ldw=None, wsj=0, wsc=1,
               pscale=None, pshift=None, pslope=1.0, dbias=None, dw_off=0):
    g = ConvWgrad()
    g.A = _p(_f32(A)); g.lda = Cin if lda is None else lda
    g.G = _p(_f32(G)); g.ldg = N if ldg is None else ldg
    g.dW = _p(_f32(dW)) + 4 * dw_off; g.ldw = taps * Cin if ldw is None else ldw; g.wsj = wsj; g.wsc = wsc
    g.B, g.Tout, g.Tin, g.N, g.Cin, g.taps, g.stride, g.dil, g.pad = B, Tout, Tin, N, Cin, taps, stride, dil, pad
    g.pscale = _p(pscale); g.pshift = _p(pshift); g.pslope = pslope
    g.dbias = _p(dbias)
    check(_L().tg_conv_wgrad_f32(ctypes.byref(g), _s()), 'tg_conv_wgrad_f32')
    _count()


def gemm_tf32(A, Bw, C, *, M, N, K, lda=None, ldb=None, ldc=None, a_rows=None, taps=1, shift0=0, T=1, escale=None, bias=None, act1=0,
              slope1=0.0, mask=None, ldmask=None, residual=None, ldres=None, act2=0, accumulate=False, clip_rows=0, a_clip_pitch=0):
    """C[M,N] = epi(A[M,K] @ Bw[N,K]^T) on the tcgen05 tensor cores (TF32 operands, fp32 accumulate); see tg_gemm_tf32_t."""
    g = GemmTf32()
    g.A = _p(_f32(A)); g.lda = K if lda is None else lda; g.a_rows = M if a_rows is None else a_rows
    g.Bw = _p(_f32(Bw)); g.ldb = K if ldb is None else ldb
    g.C = _p(_f32(C)); g.ldc = N if ldc is None else ldc
    g.M, g.N, g.K, g.taps, g.shift0, g.T = M, N, K, taps, shift0, T
    g.escale = _p(escale); g.bias = _p(bias); g.act1 = act1; g.slope1 = slope1
    g.mask = _p(mask); g.ldmask = N if ldmask is None else ldmask
    g.residual = _p(residual); g.ldres = N if ldres is None else ldres
    g.act2 = act2; g.accumulate = 1 if accumulate else 0
    g.clip_rows = clip_rows; g.a_clip_pitch = a_clip_pitch
    check(_L().tg_gemm_tf32(ctypes.byref(g), _s()), 'tg_gemm_tf32')
    _count()


def wgrad_tf32(G, X, dW, *, B, T, N, Cin, shift=0, ldg=None, ldx=None, ldw=None, dbias=None, x_clip_pitch=0):
    """dW[N,Cin] += G^T X over B clips x T rows on the tensor cores (TF32); see tg_wgrad_tf32_t."""
    g = WgradTf32()
    g.G = _p(_f32(G)); g.ldg = N if ldg is None else ldg
    g.X = _p(_f32(X)); g.ldx = Cin if ldx is None else ldx
    g.dW = _p(_f32(dW)); g.ldw = Cin if ldw is None else ldw
    g.dbias = _p(dbias)
    g.B, g.T, g.N, g.Cin, g.shift = B, T, N, Cin, shift
    g.x_clip_pitch = x_clip_pitch
    check(_L().tg_wgrad_tf32(ctypes.byref(g), _s()), 'tg_wgrad_tf32')
    _count(2 if dbias is not None else 1)


def linear(x, W, bias, out, *, M, K, N, **kw):
    """out[M,N] = x[M,K] @ W[N,K]^T + bias (nn.Linear)."""
    conv_gemm(x, W, out, B=1, Tin=M, Tout=M, N=N, Cin=K, taps=1, ldw=K, wsc=1, bias=bias, **kw)


def linear_dgrad(dy, W, dx, *, M, K, N, **kw):
    """dx[M,K] = dy[M,N] @ W[N,K]."""
    conv_gemm(dy, W, dx, B=1, Tin=M, Tout=M, N=K, Cin=N, taps=1, ldw=1, wsc=K, **kw)


def linear_wgrad(x, dy, dW, dbias, *, M, K, N, **kw):
    conv_wgrad(x, dy, dW, B=1, Tin=M, Tout=M, N=N, Cin=K, taps=1, ldw=K, wsc=1, dbias=dbias, **kw)


def conv1d(x, W, bias, y, *, B, Tin, Cin, N, k, stride=1, dil=1, pad=0, **kw):
    """nn.Conv1d on channels-last x [B,Tin,Cin] with the reference weight layout W [N,Cin,k]; y [B,Tout,N]."""
    Tout = (Tin + 2 * pad - dil * (k - 1) - 1) // stride + 1 if 'Tout' not in kw else kw.pop('Tout')
    conv_gemm(x, W, y, B=B, Tin=Tin, Tout=Tout, N=N, Cin=Cin, taps=k, stride=stride, dil=dil, pad=pad, ldw=Cin * k, wsj=1, wsc=k,
              bias=bias, **kw)
    return Tout


def conv1d_wgrad(x, dy, dW, dbias, *, B, Tin, Tout, Cin, N, k, stride=1, dil=1, pad=0, **kw):
    conv_wgrad(x, dy, dW, B=B, Tin=Tin, Tout=Tout, N=N, Cin=Cin, taps=k, stride=stride, dil=dil, pad=pad, ldw=Cin * k, wsj=1, wsc=k,
               dbias=dbias, **kw)


def conv1d_dgrad(dy, W, dx, *, B, Tin, Tout, Cin, N, k, stride=1, dil=1, pad=0, **kw):
    """dx [B,Tin,Cin] = conv-transpose of dy [B,Tout,N] with W [N,Cin,k]; `stride` dense phases for a strided conv."""
    if stride == 1:
        conv_gemm(dy, W, dx, B=B, Tin=Tout, Tout=Tin, N=Cin, Cin=N, taps=k, stride=1, dil=-dil, pad=-pad, ldw=k, wsj=1,
                  wsc=Cin * k, **kw)
        return
    assert dil == 1
    for r in range(min(stride, Tin)):
        j0 = (r + pad) % stride
        q_r = (Tin - r + stride - 1) // stride
        if j0 >= k:
            # no tap reaches these positions: gradient is zero
            dx.view(B, Tin, Cin)[:, r::stride].zero_()
            continue
        taps_r = (k - j0 + stride - 1) // stride
        off = (r + pad - j0) // stride
        conv_gemm(dy, W, dx, B=B, Tin=Tout, Tout=q_r, N=Cin, Cin=N, taps=taps_r, stride=1, dil=-1, pad=-off, ldw=k, wsj=stride,
                  wsc=Cin * k, w_off=j0, ToutFull=Tin, ostride=stride, ooff=r, **kw)


def conv1_direct(x, w, bias, y, *, B, Tin, Tout, N, taps, stride, pad):
    check(_L().tg_conv1_direct_f32(_p(x), _p(w), _p(bias), _p(y), B, Tin, Tout, N, taps, stride, pad, _s()), 'tg_conv1_direct_f32')
    _count()


# ------------------------------------------------------------------------------------------ tensor-core WavEncoder helpers
def window_weights(w, w2, w2t, N, Cin, k):
    check(_L().tg_window_weights(_p(_f32(w)), _p(_f32(w2)), _p(w2t), N, Cin, k, _s()), 'tg_window_weights'); _count()


def window_wgrad_add(dw2, dw, N, Cin, k):
    check(_L().tg_window_wgrad_add(_p(_f32(dw2)), _p(_f32(dw)), N, Cin, k, _s()), 'tg_window_wgrad_add'); _count()


def window_dgrad_weights(w, wd, N, Cin, k, stride):
    check(_L().tg_window_dgrad_weights(_p(_f32(w)), _p(_f32(wd)), N, Cin, k, stride, _s()), 'tg_window_dgrad_weights'); _count()


def conv_dgrad_tf32(dy, wd, da, *, B, Tin, Tout, Cin, N, k, stride):
    check(_L().tg_conv_dgrad_tf32(_p(_f32(dy)), _p(_f32(wd)), _p(_f32(da)), B, Tin, Tout, Cin, N, k, stride, _s()), 'tg_conv_dgrad_tf32'); _count()


def col2im(col, da, *, B, Tin, Tout, Cin, k, stride):
    check(_L().tg_col2im(_p(_f32(col)), _p(_f32(da)), B, Tin, Tout, Cin, k, stride, _s()), 'tg_col2im'); _count()


def conv1_wgrad(x, dy, dW, dbias, *, B, Tin, Tout, N, taps, stride, pad):
    check(_L().tg_conv1_wgrad(_p(_f32(x)), _p(_f32(dy)), _p(_f32(dW)), _p(dbias), B, Tin, Tout, N, taps, stride, pad, _s()), 'tg_conv1_wgrad')
    _count()


# ------------------------------------------------------------------------------------------ batch norm
def col_stats(x, ld, M, C, sums):
    check(_L().tg_col_stats_f64(_p(x), ld, M, C, _p(sums), _s()), 'tg_col_stats_f64'); _count()


def bn_finalize(sums, M, C, eps, momentum, n_updates, gamma, beta, rm, rv, nbt, mean, rstd, scale, shift):
    check(_L().tg_bn_finalize(_p(sums), M, C, eps, momentum, n_updates, _p(gamma), _p(beta), _p(rm), _p(rv), _p(nbt), _p(mean),
                              _p(rstd), _p(scale), _p(shift), _s()), 'tg_bn_finalize'); _count()


def bn_eval_fold(gamma, beta, rm, rv, eps, conv_bias, scale, shift, C):
    check(_L().tg_bn_eval_fold(_p(gamma), _p(beta), _p(rm), _p(rv), eps, _p(conv_bias), _p(scale), _p(shift), C, _s()),
          'tg_bn_eval_fold'); _count()


def affine_lrelu(x, y, M, C, scale, shift, slope):
    check(_L().tg_affine_lrelu(_p(x), _p(y), M, C, _p(scale), _p(shift), slope, _s()), 'tg_affine_lrelu'); _count()


def bn_bwd_reduce(dy, x, M, C, mean, rstd, scale, shift, slope, sums):
    check(_L().tg_bn_bwd_reduce(_p(dy), _p(x), M, C, _p(mean), _p(rstd), _p(scale), _p(shift), slope, _p(sums), _s()),
          'tg_bn_bwd_reduce'); _count()


def bn_bwd_apply(dy, x, dx, M, C, mean, rstd, scale, shift, slope, gamma, sums, dgamma, dbeta):
    check(_L().tg_bn_bwd_apply(_p(dy), _p(x), _p(dx), M, C, _p(mean), _p(rstd), _p(scale), _p(shift), slope, _p(gamma), _p(sums),
                               _p(dgamma), _p(dbeta), _s()), 'tg_bn_bwd_apply'); _count()


# ------------------------------------------------------------------------------------------ embedding / weight norm / misc
def embedding_gather(table, idx, idx_mod, mask, out, M, E):
    assert idx.dtype == torch.int64
    check(_L().tg_embedding_gather(_p(table), _p(idx), idx_mod, _p(mask), _p(out), M, E, _s()), 'tg_embedding_gather'); _count()


def embedding_scatter_add(dout, idx, mask, dtable, M, E):
    assert idx.dtype == torch.int64
    check(_L().tg_embedding_scatter_add(_p(dout), _p(idx), _p(mask), _p(dtable), M, E, _s()), 'tg_embedding_scatter_add'); _count()


def weight_norm_fwd(v, g, w, wT, inv_norm, N, Cin, taps):
    check(_L().tg_weight_norm_fwd(_p(v), _p(g), _p(w), _p(wT), _p(inv_norm), N, Cin, taps, _s()), 'tg_weight_norm_fwd'); _count()


def weight_norm_bwd(dw, v, g, inv_norm, dv, dg, N, Cin, taps):
    check(_L().tg_weight_norm_bwd(_p(dw), _p(v), _p(g), _p(inv_norm), _p(dv), _p(dg), N, Cin, taps, _s()), 'tg_weight_norm_bwd'); _count()


def mul(a, b, out, n):
    check(_L().tg_mul(_p(a), _p(b), _p(out), n, _s()), 'tg_mul'); _count()


def add(a, b, out, n, relu=False):
    check(_L().tg_add(_p(a), _p(b), _p(out), n, 1 if relu else 0, _s()), 'tg_add'); _count()


def relu_mask_bwd(dy, y, mask, dx, n):
    check(_L().tg_relu_mask_bwd(_p(dy), _p(y), _p(mask), _p(dx), n, _s()), 'tg_relu_mask_bwd'); _count()


def tcn_res_bwd(dxo, xo, x, mask, dpre, dc2, n):
    check(_L().tg_tcn_res_bwd(_p(dxo), _p(xo), _p(x), _p(mask), _p(dpre), _p(dc2), n, _s()), 'tg_tcn_res_bwd'); _count()


def sum_halves(x, out, M, H):
    check(_L().tg_sum_halves(_p(x), _p(out), M, H, _s()), 'tg_sum_halves'); _count()


def dup_halves(d, dx, M, H):
    check(_L().tg_dup_halves(_p(d), _p(dx), M, H, _s()), 'tg_dup_halves'); _count()


def reparam_fwd(mu, logvar, eps, z, n):
    check(_L().tg_reparam_fwd(_p(mu), _p(logvar), _p(eps), _p(z), n, _s()), 'tg_reparam_fwd'); _count()


def reparam_bwd(dz, logvar, eps, dmu, dlogvar, n):
    check(_L().tg_reparam_bwd(_p(dz), _p(logvar), _p(eps), _p(dmu), _p(dlogvar), n, _s()), 'tg_reparam_bwd'); _count()


def make_pre_seq(target, pre, B, T, D, n_pre):
    check(_L().tg_make_pre_seq(_p(target), _p(pre), B, T, D, n_pre, _s()), 'tg_make_pre_seq'); _count()


def gru_input_concat(pre, audio, text, z, out, B, Ba, T, Dp, Da, Dt, Dz):
    check(_L().tg_gru_input_concat(_p(pre), _p(audio), _p(text), _p(z), _p(out), B, Ba, T, Dp, Da, Dt, Dz, _s()),
          'tg_gru_input_concat'); _count()


def gru_input_split_bwd(din, daudio, dtext, dz, B, T, Dp, Da, Dt, Dz):
    check(_L().tg_gru_input_split_bwd(_p(din), _p(daudio), _p(dtext), _p(dz), B, T, Dp, Da, Dt, Dz, _s()),
          'tg_gru_input_split_bwd'); _count()


def transpose(x, out, R, C):
    check(_L().tg_transpose_f32(_p(x), _p(out), R, C, _s()), 'tg_transpose_f32'); _count()


# ------------------------------------------------------------------------------------------ GRU
def gru_sync_ints(B, H):
    return _L().tg_gru_sync_ints(B, H)


def gru_bwd_scratch_floats(B, H):
    return _L().tg_gru_bwd_scratch_floats(B, H)


def gru_layer_fwd(gi, whhT_f, whhT_r, bhh_f, bhh_r, out, saved, saved_qstride, sync, B, T, H):
    check(_L().tg_gru_layer_fwd(_p(gi), _p(whhT_f), _p(whhT_r), _p(bhh_f), _p(bhh_r), _p(out), _p(saved), saved_qstride, _p(sync),
                                B, T, H, _s()), 'tg_gru_layer_fwd'); _count(2)


def gru_layer_bwd(dout, out, saved, saved_qstride, whh_f, whh_r, dgi, dgh, partial, sync, B, T, H):
    check(_L().tg_gru_layer_bwd(_p(dout), _p(out), _p(saved), saved_qstride, _p(whh_f), _p(whh_r), _p(dgi), _p(dgh), _p(partial),
                                _p(sync), B, T, H, _s()), 'tg_gru_layer_bwd'); _count(2)


def gru_tf32_sync_ints(B, H):
    return _L().tg_gru_tf32_sync_ints(B, H)


def gru_bwd_tf32_scratch_floats(B, H):
    return _L().tg_gru_bwd_tf32_scratch_floats(B, H)


def gru_layer_fwd_tf32(gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, saved_qstride, sync, B, T, H):
    check(_L().tg_gru_layer_fwd_tf32(_p(gi), _p(whh_f), _p(whh_r), _p(bhh_f), _p(bhh_r), _p(out), _p(saved), saved_qstride, _p(sync),
                                     B, T, H, _s()), 'tg_gru_layer_fwd_tf32'); _count(2)


def col_sum(g, ld, M, N, out):
    """out[n] += sum_m g[m*ld + n]"""
    check(_L().tg_col_sum_f32(_p(g), ld, M, N, _p(out), _s()), 'tg_col_sum_f32'); _count()


def gru_layer_fwd_tf32_drop(gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, saved_qstride, mask, drop, sync, B, T, H):
    """tg_gru_layer_fwd_tf32 + drop = out * mask written by the recurrence kernel itself"""
    check(_L().tg_gru_layer_fwd_tf32_drop(_p(gi), _p(whh_f), _p(whh_r), _p(bhh_f), _p(bhh_r), _p(out), _p(saved), saved_qstride, _p(mask),
                                          _p(drop), _p(sync), B, T, H, _s()), 'tg_gru_layer_fwd_tf32_drop'); _count(2)


def gru_layer_bwd_tf32(dout, out, saved, saved_qstride, whhT_f, whhT_r, dgi, dgh, partial, sync, B, T, H):
    check(_L().tg_gru_layer_bwd_tf32(_p(dout), _p(out), _p(saved), saved_qstride, _p(whhT_f), _p(whhT_r), _p(dgi), _p(dgh), _p(partial),
                                     _p(sync), B, T, H, _s()), 'tg_gru_layer_bwd_tf32'); _count(2)


def _ptr_array(tensors, n):
    """ctypes array of n device pointers (None -> NULL); the C side copies it into the kernel's parameter block."""
    import ctypes
    arr = (ctypes.c_void_p * n)()
    for i in range(n):
        t = tensors[i] if (tensors is not None and i < len(tensors)) else None
        arr[i] = None if t is None else t.data_ptr()
    return arr


def dgru_stack_fwd(x, gru_params, masks, outs, saved, saved_qstride, drops, w_out, b_out, w_out2, b_out2, hsum, o1, prob, B, T, I0, H, L,
                   fast=False):
    import ctypes
    a_m = _ptr_array(masks, 4) if masks is not None else None
    a_o = _ptr_array(outs, 4)
    a_s = _ptr_array(saved, 4) if saved is not None else None
    a_d = _ptr_array(drops, 4) if drops is not None else None
    addr = lambda a: None if a is None else ctypes.cast(a, ctypes.c_void_p)
    check(_L().tg_dgru_stack_fwd(_p(x), _p(gru_params), addr(a_m), addr(a_o), addr(a_s), saved_qstride, addr(a_d), _p(w_out), _p(b_out),
                                 _p(w_out2), _p(b_out2), _p(hsum), _p(o1), _p(prob), B, T, I0, H, L, 1 if fast else 0, _s()), 'tg_dgru_stack_fwd'); _count()


def dgru_stack_bwd(dlogit, gru_params, masks, outs, saved, saved_qstride, hsum, o1, w_out, w_out2, dgi, dgh, dx0, g_w_out, g_b_out, g_w_out2,
                   g_b_out2, B, T, I0, H, L, fast=False):
    import ctypes
    addr = lambda a: None if a is None else ctypes.cast(a, ctypes.c_void_p)
    a_m = _ptr_array(masks, 4) if masks is not None else None
    a_o, a_s, a_gi, a_gh = _ptr_array(outs, 4), _ptr_array(saved, 4), _ptr_array(dgi, 4), _ptr_array(dgh, 4)
    check(_L().tg_dgru_stack_bwd(_p(dlogit), _p(gru_params), addr(a_m), addr(a_o), addr(a_s), saved_qstride, _p(hsum), _p(o1), _p(w_out),
                                 _p(w_out2), addr(a_gi), addr(a_gh), _p(dx0), _p(g_w_out), _p(g_b_out), _p(g_w_out2), _p(g_b_out2),
                                 B, T, I0, H, L, 1 if fast else 0, _s()), 'tg_dgru_stack_bwd'); _count()


def dconv_stack_fwd(x, w1, b1, g1, be1, rm1, rv1, nbt1, w2, b2, g2, be2, rm2, rv2, nbt2, w3, b3, y0, y1, y2, st1, st2, B, T, D, training, eps,
                    momentum):
    check(_L().tg_dconv_stack_fwd(_p(x), _p(w1), _p(b1), _p(g1), _p(be1), _p(rm1), _p(rv1), _p(nbt1), _p(w2), _p(b2), _p(g2), _p(be2), _p(rm2),
                                  _p(rv2), _p(nbt2), _p(w3), _p(b3), _p(y0), _p(y1), _p(y2), _p(st1), _p(st2), B, T, D, int(training), eps,
                                  momentum, _s()), 'tg_dconv_stack_fwd'); _count()


def dconv_stack_bwd(dy2, x, y0, y1, st1, st2, w1, w2, w3, g1, g2, dw1, db1, dw2, db2, dw3, db3, dg1, dbe1, dg2, dbe2, dx, B, T, D):
    check(_L().tg_dconv_stack_bwd(_p(dy2), _p(x), _p(y0), _p(y1), _p(st1), _p(st2), _p(w1), _p(w2), _p(w3), _p(g1), _p(g2), _p(dw1), _p(db1),
                                  _p(dw2), _p(db2), _p(dw3), _p(db3), _p(dg1), _p(dbe1), _p(dg2), _p(dbe2), _p(dx), B, T, D, _s()),
          'tg_dconv_stack_bwd'); _count()


# ------------------------------------------------------------------------------------------ losses / optimiser / rng
def gen_losses(out, target, out_rand, z, z_rand, mu, logvar, B, TD, Z, w_reg, w_div, w_kld, scalars, d_out, dmu, dlogvar):
    check(_L().tg_gen_losses(_p(out), _p(target), _p(out_rand), _p(z), _p(z_rand), _p(mu), _p(logvar), B, TD, Z, w_reg, w_div, w_kld,
                             _p(scalars), _p(d_out), _p(dmu), _p(dlogvar), _s()), 'tg_gen_losses'); _count()


def bce_sigmoid(p, n, s, o, w, scalar, dlogit):
    check(_L().tg_bce_sigmoid(_p(p), n, s, o, w, _p(scalar), _p(dlogit), _s()), 'tg_bce_sigmoid'); _count()


def adam_flat(p, g, m, v, n, lr, b1, b2, eps, grad_scale, step_dev):
    check(_L().tg_adam_flat(_p(p), _p(g), _p(m), _p(v), n, lr, b1, b2, eps, grad_scale, _p(step_dev), _s()), 'tg_adam_flat'); _count()


def increment_i64(x, by=1):
    check(_L().tg_increment_i64(_p(x), by, _s()), 'tg_increment_i64'); _count()


def philox_normal(out, n, seed, offset_dev, stream_id):
    check(_L().tg_philox_normal(_p(out), n, seed, _p(offset_dev), stream_id, _s()), 'tg_philox_normal'); _count()


def philox_dropout_mask(out, n, p, seed, offset_dev, stream_id):
    check(_L().tg_philox_dropout_mask(_p(out), n, p, seed, _p(offset_dev), stream_id, _s()), 'tg_philox_dropout_mask'); _count()


def philox_randperm(out, n, seed, offset_dev, stream_id):
    check(_L().tg_philox_randperm(_p(out), n, seed, _p(offset_dev), stream_id, _s()), 'tg_philox_randperm'); _count()


def gather_i64(src, idx, out, n):
    check(_L().tg_gather_i64(_p(src), _p(idx), _p(out), n, _s()), 'tg_gather_i64'); _count()


def feature_stats(feat, n, F, acc):
    check(_L().tg_feature_stats_f64(_p(feat), n, F, _p(acc), _s()), 'tg_feature_stats_f64'); _count()


def l1_dist(a, b, n, acc):
    check(_L().tg_l1_dist_f64(_p(a), _p(b), n, _p(acc), _s()), 'tg_l1_dist_f64'); _count()


# ------------------------------------------------------------------------------------------ seq2seq baseline
def gru_gates_fwd(gi, ldgi, gh, hprev, lengths, t, hnew, out, ldout, saved, saved_plane, B, H):
    check(_L().tg_gru_gates_fwd(_p(gi), ldgi, _p(gh), _p(hprev), _p(lengths), t, _p(hnew), _p(out), ldout, _p(saved), saved_plane, B, H, _s()),
          'tg_gru_gates_fwd'); _count()


def gru_gates_bwd(dh, dadd, ldadd, saved, saved_plane, hprev, lengths, t, dgi, lddgi, dgh, dhprev, B, H):
    check(_L().tg_gru_gates_bwd(_p(dh), _p(dadd), ldadd, _p(saved), saved_plane, _p(hprev), _p(lengths), t, _p(dgi), lddgi, _p(dgh),
                                _p(dhprev), B, H, _s()), 'tg_gru_gates_bwd'); _count()


def attn_fwd(hq, eproj, enc, v, w, ctx, B, Tm, H):
    check(_L().tg_attn_fwd(_p(hq), _p(eproj), _p(enc), _p(v), _p(w), _p(ctx), B, Tm, H, _s()), 'tg_attn_fwd'); _count()


def attn_bwd(dctx, w, hq, eproj, enc, v, denc, deproj, dv, dhq, B, Tm, H):
    check(_L().tg_attn_bwd(_p(dctx), _p(w), _p(hq), _p(eproj), _p(enc), _p(v), _p(denc), _p(deproj), _p(dv), _p(dhq), B, Tm, H, _s()),
          'tg_attn_bwd'); _count()


def s2s_loss(out, target, loss, dy_tmajor, B, T, D, w_mse, w_cont, w_var):
    check(_L().tg_s2s_loss(_p(out), _p(target), _p(loss), _p(dy_tmajor), B, T, D, float(w_mse), float(w_cont), float(w_var), _s()),
          'tg_s2s_loss'); _count()


def s2s_gather_inputs(poses, outputs, xin, B, T, D, n_pre):
    check(_L().tg_s2s_gather_inputs(_p(poses), _p(outputs), _p(xin), B, T, D, n_pre, _s()), 'tg_s2s_gather_inputs'); _count()


def sumsq(x, n, out):
    check(_L().tg_sumsq_f64(_p(x), n, _p(out), _s()), 'tg_sumsq_f64'); _count()


def clip_scale(x, n, sumsq_dev, max_norm):
    check(_L().tg_clip_scale(_p(x), n, _p(sumsq_dev), float(max_norm), _s()), 'tg_clip_scale'); _count()


def pose_eval_metrics(out, target, B, T, D, n_pre, acc):
    check(_L().tg_pose_eval_metrics(_p(_f32(out)), _p(_f32(target)), B, T, D, n_pre, _p(acc), _s()), 'tg_pose_eval_metrics'); _count()


def ae_recon_loss(recon, target, B, T, D, use_diff, weight, acc, d_recon):
    check(_L().tg_ae_recon_loss(_p(_f32(recon)), _p(_f32(target)), B, T, D, 1 if use_diff else 0, float(weight), _p(acc), _p(d_recon), _s()),
          'tg_ae_recon_loss'); _count()


def transpose_batched(x, out, B, R, C):
    """out[b][c][r] = x[b][r][c]"""
    check(_L().tg_transpose_batched_f32(_p(_f32(x)), _p(_f32(out)), B, R, C, _s()), 'tg_transpose_batched_f32'); _count()


def copy_bytes(dst, src, max_ctas=64):
    """dst (device tensor) <- src (device or PINNED host tensor, same byte size), by a kernel (no copy engine); see tg_copy_bytes."""
    assert dst.is_cuda and dst.is_contiguous() and src.is_contiguous() and (src.is_cuda or src.is_pinned())
    nbytes = dst.numel() * dst.element_size()
    assert nbytes == src.numel() * src.element_size()
    check(_L().tg_copy_bytes(dst.data_ptr(), src.data_ptr(), nbytes, max_ctas, _s()), 'tg_copy_bytes'); _count()


# ---- Speech2Gesture pieces (csrc/s2g.cu)
def im2col2d(x, col, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo):
    check(_L().tg_im2col2d(_p(x), _p(col), B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo, _s()), 'tg_im2col2d'); _count()


def col2im2d(col, dx, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo):
    check(_L().tg_col2im2d(_p(col), _p(dx), B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo, _s()), 'tg_col2im2d'); _count()


def resize_bilinear_fwd(x, y, B, H, W, C, Ho, Wo):
    check(_L().tg_resize_bilinear_fwd(_p(x), _p(y), B, H, W, C, Ho, Wo, _s()), 'tg_resize_bilinear_fwd'); _count()


def resize_bilinear_bwd(dy, dx, B, H, W, C, Ho, Wo):
    check(_L().tg_resize_bilinear_bwd(_p(dy), _p(dx), B, H, W, C, Ho, Wo, _s()), 'tg_resize_bilinear_bwd'); _count()


def upsample2_add_fwd(x1, x2, y, B, T1, T2, C):
    check(_L().tg_upsample2_add_fwd(_p(x1), _p(x2), _p(y), B, T1, T2, C, _s()), 'tg_upsample2_add_fwd'); _count()


def upsample2_bwd(dy, dx1, B, T1, T2, C, accumulate=False):
    check(_L().tg_upsample2_bwd(_p(dy), _p(dx1), B, T1, T2, C, 1 if accumulate else 0, _s()), 'tg_upsample2_bwd'); _count()


def time_diff_fwd(x, y, B, T, D):
    check(_L().tg_time_diff_fwd(_p(x), _p(y), B, T, D, _s()), 'tg_time_diff_fwd'); _count()


def time_diff_bwd(dy, dx, B, T, D, accumulate=False):
    check(_L().tg_time_diff_bwd(_p(dy), _p(dx), B, T, D, 1 if accumulate else 0, _s()), 'tg_time_diff_bwd'); _count()


def concat_bcast_fwd(a, p, y, B, T, Ca, Cp):
    check(_L().tg_concat_bcast_fwd(_p(a), _p(p), _p(y), B, T, Ca, Cp, _s()), 'tg_concat_bcast_fwd'); _count()


def concat_bcast_bwd(d, da, dp, B, T, Ca, Cp):
    check(_L().tg_concat_bcast_bwd(_p(d), _p(da), _p(dp), B, T, Ca, Cp, _s()), 'tg_concat_bcast_bwd'); _count()


def lrelu_bwd(dy, x, dx, n, slope):
    check(_L().tg_lrelu_bwd(_p(dy), _p(x), _p(dx), n, float(slope), _s()), 'tg_lrelu_bwd'); _count()


def mse_const(x, n, target, w, scalar, dx):
    check(_L().tg_mse_const(_p(x), n, float(target), float(w), _p(scalar), _p(dx), _s()), 'tg_mse_const'); _count()


def l1_loss(x, y, n, w, scalar, dx):
    check(_L().tg_l1_loss(_p(x), _p(y), n, float(w), _p(scalar), _p(dx), _s()), 'tg_l1_loss'); _count()
