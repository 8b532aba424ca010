"""Launch plan of the seq2seq baseline (scripts/model/seq2seq_net.py:14-254, scripts/train_eval/train_seq2seq.py:6-51):
packed bidirectional GRU encoder, Bahdanau-attention GRU decoder unrolled over the 33 generated frames, custom_loss,
hand-derived backward through time, global-norm clipping and flat Adam - every FLOP in the C-ABI kernels (tgb200.ops).

Layout.  Encoder tensors are batch-major [B, Tm, .] (one GEMM projects all time steps); everything the decoder saves per
step is TIME-major [T, B, .] so a step's slice is contiguous and the weight gradients of the whole unrolled decoder are
one GEMM per weight over the stacked [(T-1)*B, .] rows after the backward sweep (instead of 33 small ones).
What the reference recomputes every step but does not change is hoisted: the encoder half of the attention projection
(W_a[:, H:] enc) is computed once per forward, not once per step (seq2seq_net.py:87 concatenates and multiplies 33 times)."""
from __future__ import annotations

from typing import Optional

import torch

from . import config, ops
from .arena import ParamArena
from .engine import BN_EPS, BN_MOM, Workspace

_WT = {}          # id(weight tensor storage) -> transposed copy [K_total, N] of the current iteration (fast mode, see prep_weights)


def _al16(t):
    return t.data_ptr() % 16 == 0


def _lin(x, W, b, out, *, M, K, N, ldw=None, w_off=0, lda=None, ldc=None, accumulate=False):
    """out[M,N] (+)= x[M,K] @ W[:, w_off:w_off+K]^T + b   (W rows have pitch ldw).  Fast mode: tcgen05 TF32 GEMM when TMA can
    describe the operands (16-byte aligned bases and pitches), fp32 FFMA GEMM otherwise."""
    ldw_ = K if ldw is None else ldw
    lda_ = K if lda is None else lda
    if (config.fast() and K >= 8 and K % 4 == 0 and lda_ % 4 == 0 and ldw_ % 4 == 0 and w_off % 4 == 0 and _al16(x) and _al16(W)):
        ops.gemm_tf32(x, W.reshape(-1)[w_off:], out, M=M, N=N, K=K, lda=lda_, ldb=ldw_, ldc=ldc, bias=b, accumulate=accumulate)
    else:
        ops.conv_gemm(x, W, out, B=1, Tin=M, Tout=M, N=N, Cin=K, taps=1, ldw=ldw_, wsc=1, w_off=w_off, lda=lda, ldc=ldc, bias=b,
                      accumulate=accumulate)


def _dlin(dy, W, dx, *, M, K, N, ldw=None, w_off=0, ldc=None, accumulate=False):
    """dx[M,K] (+)= dy[M,N] @ W[:, w_off:w_off+K].  Fast mode reads the transposed copy W^T prepared by prep_weights()."""
    ldw_ = K if ldw is None else ldw
    wt = _WT.get(W.data_ptr()) if config.fast() else None
    if (wt is not None and N >= 8 and N % 4 == 0 and (w_off * N) % 4 == 0 and _al16(dy)):
        ops.gemm_tf32(dy, wt.reshape(-1)[w_off * N:], dx, M=M, N=K, K=N, ldc=ldc, accumulate=accumulate)
    else:
        ops.conv_gemm(dy, W, dx, B=1, Tin=M, Tout=M, N=K, Cin=N, taps=1, ldw=1, wsc=ldw_, w_off=w_off, ldc=ldc, accumulate=accumulate)


def _wg(x, dy, dW, dbias, *, M, K, N, ldw=None, dw_off=0):
    """dW[:, dw_off:dw_off+K] += dy[M,N]^T x[M,K];  dbias += colsum(dy)"""
    ldw_ = K if ldw is None else ldw
    if (config.fast() and K >= 8 and K % 4 == 0 and N >= 16 and N % 4 == 0 and ldw_ % 4 == 0 and dw_off % 4 == 0 and _al16(x) and _al16(dy)
            and _al16(dW)):
        ops.wgrad_tf32(dy, x, dW.reshape(-1)[dw_off:], B=1, T=M, N=N, Cin=K, ldw=ldw_, dbias=dbias)
    else:
        ops.conv_wgrad(x, dy, dW, B=1, Tin=M, Tout=M, N=N, Cin=K, taps=1, ldw=ldw_, wsc=1, dbias=dbias, dw_off=dw_off)


class Seq2SeqEngine:
    def __init__(self, module):
        self.m = module
        self.arena = ParamArena(module)
        self.H = module.encoder.hidden_size
        self.L = module.encoder.n_layers
        self.E = module.encoder.embed_size
        self.D = module.decoder.output_size
        self.T = module.n_frames
        self.n_pre = module.n_pre_poses
        self.p_enc = float(module.encoder.dropout)
        self.p_dec = float(module.decoder.decoder.dropout_p)
        assert module.decoder.decoder.n_layers == self.L
        assert not module.decoder.decoder.discrete_representation and module.decoder.decoder.speaker_model is None, \
            'discrete / speaker-conditioned decoders are not on the configured path (config/seq2seq.yml)'
        self.ws: Optional[Workspace] = None
        self.slots = {}

    def ensure(self, device, slot='default'):
        if not self.arena.is_current():
            self.slots = {}
        self.arena.ensure(device)
        key = (slot, str(device))
        if key not in self.slots:
            self.slots[key] = Workspace(device)
        self.ws = self.slots[key]
        self.bufs = dict(self.m.named_buffers())
        return self

    def P(self, name):
        return self.arena.params[name].data

    def G(self, name):
        return self.arena.gview(name)

    def prep_weights(self):
        """Fast mode: transposed copies of every matrix the backward data path multiplies by (weights change every optimiser step)."""
        if not config.fast():
            return
        ws = self.ws
        names = [n for n in self.arena.names if '.gru.weight_' in n or n.endswith(('attn.attn.weight', 'pre_linear.0.weight'))]
        for n in dict.fromkeys(names):
            W = self.P(n)
            if W.dim() != 2 or W.shape[0] % 4 != 0:
                continue
            wt = ws.get('T.' + n, (W.shape[1], W.shape[0]))
            ops.transpose(W, wt, W.shape[0], W.shape[1])
            _WT[W.data_ptr()] = wt

    # ------------------------------------------------------------------------------------------------ masks
    def make_masks(self, B, Tm, seed, offset_dev):
        """Inter-layer dropout keep-masks (train mode): encoder layer l < L-1 outputs [B*Tm, 2H], decoder layer l < L-1 outputs of
        every step [T, B, H] (nn.GRU dropout=, seq2seq_net.py:31,129)."""
        ws, H, L, T = self.ws, self.H, self.L, self.T
        masks = {}
        sid = 0
        for l in range(L - 1):
            if self.p_enc > 0:
                mk = ws.get(f'mask.enc{l}', (B * Tm, 2 * H))
                ops.philox_dropout_mask(mk, B * Tm * 2 * H, self.p_enc, seed, offset_dev, sid)
                masks[f'enc{l}'] = mk
            sid += 1
        for l in range(L - 1):
            if self.p_dec > 0:
                mk = ws.get(f'mask.dec{l}', (T, B, H))
                ops.philox_dropout_mask(mk, T * B * H, self.p_dec, seed, offset_dev, sid)
                masks[f'dec{l}'] = mk
            sid += 1
        return masks

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, in_text, lengths_dev, Tm, poses, training, masks=None, save=True):
        """Seq2SeqNet.forward (seq2seq_net.py:229-254).  in_text [B, >=Tm] int64, lengths_dev [B] int64 (device), Tm = max
        length (host int), poses [B,T,D].  Returns outputs [B,T,D] (a workspace view)."""
        ws, H, L, E, D, T = self.ws, self.H, self.L, self.E, self.D, self.T
        B = poses.shape[0]
        self.ctx = dict(B=B, Tm=Tm, masks=masks, poses=poses, training=training, lengths=lengths_dev)
        if save:
            self.prep_weights()
        idx = ws.get('enc.idx', (B, Tm), torch.int64)
        idx.copy_(in_text[:, :Tm])
        emb = ws.get('enc.emb', (B * Tm, E))
        ops.embedding_gather(self.P('encoder.embedding.weight'), idx, 0, None, emb, B * Tm, E)
        # ---- encoder: 2-layer bidirectional GRU over packed sequences (length mask inside the gate kernel)
        inp, I = emb, E
        gh = ws.get('enc.gh', (B, 3 * H))
        for l in range(L):
            lo = ws.get(f'enc.out{l}', (B * Tm, 2 * H))
            for d, suf in enumerate(('', '_reverse')):
                q = f'encoder.gru.%s_l{l}{suf}'
                gi = ws.get(f'enc.gi{l}_{d}', (B * Tm, 3 * H))
                _lin(inp, self.P(q % 'weight_ih'), self.P(q % 'bias_ih'), gi, M=B * Tm, K=I, N=3 * H)
                hs = ws.get(f'enc.hs{l}_{d}', (Tm + 1, B, H))
                hs[0].zero_()
                sv = ws.get(f'enc.saved{l}_{d}', (Tm, 4, B, H)) if save else None
                for step in range(Tm):
                    t = step if d == 0 else Tm - 1 - step
                    _lin(hs[step], self.P(q % 'weight_hh'), self.P(q % 'bias_hh'), gh, M=B, K=H, N=3 * H)
                    ops.gru_gates_fwd(gi[t:], Tm * 3 * H, gh, hs[step], lengths_dev, t, hs[step + 1], lo[t:, d * H:], Tm * 2 * H,
                                      sv[step] if save else None, B * H, B, H)
            if training and masks is not None and l < L - 1 and f'enc{l}' in masks:
                dr = ws.get(f'enc.drop{l}', (B * Tm, 2 * H))
                ops.mul(lo, masks[f'enc{l}'], dr, B * Tm * 2 * H)
                inp = dr
            else:
                inp = lo
            I = 2 * H
        enc = ws.get('enc.sum', (B * Tm, H))
        ops.sum_halves(inp, enc, B * Tm, H)
        # ---- decoder
        pd = 'decoder.decoder.'
        Wa, ba, v = self.P(pd + 'attn.attn.weight'), self.P(pd + 'attn.attn.bias'), self.P(pd + 'attn.v')
        Wp, bp = self.P(pd + 'pre_linear.0.weight'), self.P(pd + 'pre_linear.0.bias')
        bn = pd + 'pre_linear.1'
        eproj = ws.get('dec.eproj', (B * Tm, H))
        _lin(enc, Wa, None, eproj, M=B * Tm, K=H, N=H, ldw=2 * H, w_off=H)            # encoder half of the attention projection, once
        outputs = ws.get('dec.outputs', (B, T, D))
        outputs[:, 0].copy_(poses[:, 0])
        Hs = [ws.get(f'dec.h{l}', (T, B, H)) for l in range(L)]                       # Hs[l][t] = hidden of layer l after step t
        for i in range(L):                                                            # encoder_hidden[:n_layers] (:241): torch order l0 fwd, l0 rev, l1 fwd, ...
            Hs[i][0].copy_(ws[f'enc.hs{i // 2}_{i % 2}'][Tm])
        hq = ws.get('dec.hq', (T, B, H)); wts = ws.get('dec.w', (T, B, Tm)); ctxs = ws.get('dec.ctx', (T, B, H))
        pre = ws.get('dec.pre', (T, B, H)); act = ws.get('dec.act', (T, B, H))
        mean = ws.get('dec.mean', (T, H)); rstd = ws.get('dec.rstd', (T, H)); scale = ws.get('dec.scale', (T, H)); shift = ws.get('dec.shift', (T, H))
        sums = ws.get('dec.sums', (2 * H,), torch.float64)
        gi = ws.get('dec.gi', (B, 3 * H)); gh = ws.get('dec.gh', (B, 3 * H))
        sv = [ws.get(f'dec.saved{l}', (T, 4, B, H)) for l in range(L)] if save else None
        xdrop = [ws.get(f'dec.xdrop{l}', (T, B, H)) for l in range(L - 1)]
        if not training:
            ops.bn_eval_fold(self.P(bn + '.weight'), self.P(bn + '.bias'), self.bufs[bn + '.running_mean'], self.bufs[bn + '.running_var'],
                             BN_EPS, None, scale[0], shift[0], H)
        TD = T * D
        for t in range(1, T):
            x_in = (poses if t - 1 < self.n_pre else outputs)[:, t - 1]              # [B, D] rows with pitch T*D
            _lin(Hs[L - 1][t - 1], Wa, ba, hq[t], M=B, K=H, N=H, ldw=2 * H)
            ops.attn_fwd(hq[t], eproj, enc, v, wts[t], ctxs[t], B, Tm, H)
            _lin(x_in, Wp, bp, pre[t], M=B, K=D, N=H, ldw=D + H, lda=TD)
            _lin(ctxs[t], Wp, None, pre[t], M=B, K=H, N=H, ldw=D + H, w_off=D, accumulate=True)
            if training:
                sums.zero_()
                ops.col_stats(pre[t], H, B, H, sums)
                ops.bn_finalize(sums, B, H, BN_EPS, BN_MOM, 1, self.P(bn + '.weight'), self.P(bn + '.bias'), self.bufs[bn + '.running_mean'],
                                self.bufs[bn + '.running_var'], self.bufs[bn + '.num_batches_tracked'], mean[t], rstd[t], scale[t], shift[t])
                sc, sh = scale[t], shift[t]
            else:
                sc, sh = scale[0], shift[0]
            ops.affine_lrelu(pre[t], act[t], B, H, sc, sh, 0.0)                        # BatchNorm + ReLU
            x = act[t]
            for l in range(L):
                q = pd + f'gru.%s_l{l}'
                _lin(x, self.P(q % 'weight_ih'), self.P(q % 'bias_ih'), gi, M=B, K=H, N=3 * H)
                _lin(Hs[l][t - 1], self.P(q % 'weight_hh'), self.P(q % 'bias_hh'), gh, M=B, K=H, N=3 * H)
                ops.gru_gates_fwd(gi, 3 * H, gh, Hs[l][t - 1], None, 0, Hs[l][t], None, 0, sv[l][t] if save else None, B * H, B, H)
                x = Hs[l][t]
                if training and masks is not None and l < L - 1 and f'dec{l}' in masks:
                    ops.mul(x, masks[f'dec{l}'][t], xdrop[l][t], B * H)
                    x = xdrop[l][t]
            _lin(x, self.P(pd + 'out.weight'), self.P(pd + 'out.bias'), outputs[:, t], M=B, K=H, N=D, ldc=TD)
        return outputs

    # ------------------------------------------------------------------------------------------------ loss + backward
    def loss_backward(self, target, w_mse, w_cont, w_var, loss_out):
        """custom_loss (train_seq2seq.py:6-36) on the last forward + backward through decoder and encoder; accumulates every
        parameter gradient into the flat gradient arena.  loss_out: fp64 [1] device accumulator (zeroed by the caller)."""
        ws, H, L, E, D, T = self.ws, self.H, self.L, self.E, self.D, self.T
        c = self.ctx
        B, Tm, masks, training, lengths = c['B'], c['Tm'], c['masks'], c['training'], c['lengths']
        assert training, 'backward through eval-mode BatchNorm is not on the reference path'
        outputs = ws['dec.outputs']
        dy = ws.get('dec.dy', (T, B, D))
        ops.s2s_loss(outputs, target, loss_out, dy, B, T, D, w_mse, w_cont, w_var)
        self.backward_from_dy(dy)

    def backward_from_dy(self, dy):
        """Backward through decoder and encoder from dy [T,B,D] = d loss / d outputs, TIME-major (row t = 0 is ignored: frame 0 is a copy
        of the input pose, seq2seq_net.py:244-245); accumulates every parameter gradient into the flat gradient arena."""
        ws, H, L, E, D, T = self.ws, self.H, self.L, self.E, self.D, self.T
        c = self.ctx
        B, Tm, masks, training, lengths = c['B'], c['Tm'], c['masks'], c['training'], c['lengths']
        assert training, 'backward through eval-mode BatchNorm is not on the reference path'
        outputs = ws['dec.outputs']
        pd = 'decoder.decoder.'
        Wa, v = self.P(pd + 'attn.attn.weight'), self.P(pd + 'attn.v')
        Wp = self.P(pd + 'pre_linear.0.weight')
        bn = pd + 'pre_linear.1'
        enc, eproj = ws['enc.sum'], ws['dec.eproj']
        Hs = [ws[f'dec.h{l}'] for l in range(L)]
        hq, wts, ctxs, pre, act = ws['dec.hq'], ws['dec.w'], ws['dec.ctx'], ws['dec.pre'], ws['dec.act']
        mean, rstd, scale, shift = ws['dec.mean'], ws['dec.rstd'], ws['dec.scale'], ws['dec.shift']
        sv = [ws[f'dec.saved{l}'] for l in range(L)]
        # stacked per-step gradients (row t = 0 stays zero: allocated zeroed, never written)
        dgi = [ws.get(f'dec.dgi{l}', (T, B, 3 * H), zero=True) for l in range(L)]
        dgh = [ws.get(f'dec.dgh{l}', (T, B, 3 * H), zero=True) for l in range(L)]
        dpre = ws.get('dec.dpre', (T, B, H), zero=True)
        dhq = ws.get('dec.dhq', (T, B, H), zero=True)
        dH = [ws.get(f'dec.dH{l}', (B, H)) for l in range(L)]
        for l in range(L):
            dH[l].zero_()
        denc = ws.get('dec.denc', (B * Tm, H)); denc.zero_()
        deproj = ws.get('dec.deproj', (B * Tm, H)); deproj.zero_()
        dx = ws.get('dec.dx', (B, H)); dact = ws.get('dec.dact', (B, H)); dctx = ws.get('dec.dctx', (B, H))
        bsums = ws.get('dec.bsums', (2 * H,), torch.float64)
        for t in range(T - 1, 0, -1):
            # out = W_out x_top
            _dlin(dy[t], self.P(pd + 'out.weight'), dH[L - 1], M=B, K=H, N=D, accumulate=True)
            dadd = None
            for l in range(L - 1, -1, -1):
                q = pd + f'gru.%s_l{l}'
                ops.gru_gates_bwd(dH[l], dadd, H, sv[l][t], B * H, Hs[l][t - 1], None, 0, dgi[l][t], 3 * H, dgh[l][t], dH[l], B, H)
                _dlin(dgh[l][t], self.P(q % 'weight_hh'), dH[l], M=B, K=H, N=3 * H, accumulate=True)
                # gradient of this layer's input: the layer below's output (through its dropout mask), or BN+ReLU for layer 0
                tgt = dx if l > 0 else dact
                _dlin(dgi[l][t], self.P(q % 'weight_ih'), tgt, M=B, K=H, N=3 * H)
                if l > 0:
                    if masks is not None and f'dec{l - 1}' in masks:
                        ops.mul(dx, masks[f'dec{l - 1}'][t], dx, B * H)
                    dadd = dx
            # BatchNorm (train mode) + ReLU
            bsums.zero_()
            ops.bn_bwd_reduce(dact, pre[t], B, H, mean[t], rstd[t], scale[t], shift[t], 0.0, bsums)
            ops.bn_bwd_apply(dact, pre[t], dpre[t], B, H, mean[t], rstd[t], scale[t], shift[t], 0.0, self.P(bn + '.weight'), bsums,
                             self.G(bn + '.weight'), self.G(bn + '.bias'))
            # pre_linear: input half feeds back into the previous step's output gradient once the decoder is free-running
            if t - 1 >= self.n_pre:
                _dlin(dpre[t], Wp, dy[t - 1], M=B, K=D, N=H, ldw=D + H, accumulate=True)
            _dlin(dpre[t], Wp, dctx, M=B, K=H, N=H, ldw=D + H, w_off=D)
            ops.attn_bwd(dctx, wts[t], hq[t], eproj, enc, v, denc, deproj, self.G(pd + 'attn.v'), dhq[t], B, Tm, H)
            _dlin(dhq[t], Wa, dH[L - 1], M=B, K=H, N=H, ldw=2 * H, accumulate=True)
        # ---- weight gradients of the unrolled decoder: one GEMM per weight over the stacked (T-1)*B rows
        M1 = (T - 1) * B
        xin = ws.get('dec.xin', (T, B, D))
        ops.s2s_gather_inputs(c['poses'], outputs, xin, B, T, D, self.n_pre)
        _wg(Hs[L - 1][1:], dy[1:], self.G(pd + 'out.weight'), self.G(pd + 'out.bias'), M=M1, K=H, N=D)
        for l in range(L):
            q = pd + f'gru.%s_l{l}'
            if l == 0:
                xl = act
            elif masks is not None and f'dec{l - 1}' in masks:
                xl = ws[f'dec.xdrop{l - 1}']
            else:
                xl = Hs[l - 1]
            _wg(xl[1:], dgi[l][1:], self.G(q % 'weight_ih'), self.G(q % 'bias_ih'), M=M1, K=H, N=3 * H)
            _wg(Hs[l][:-1], dgh[l][1:], self.G(q % 'weight_hh'), self.G(q % 'bias_hh'), M=M1, K=H, N=3 * H)
        _wg(xin[1:], dpre[1:], self.G(pd + 'pre_linear.0.weight'), self.G(pd + 'pre_linear.0.bias'), M=M1, K=D, N=H, ldw=D + H)
        _wg(ctxs[1:], dpre[1:], self.G(pd + 'pre_linear.0.weight'), None, M=M1, K=H, N=H, ldw=D + H, dw_off=D)
        _wg(Hs[L - 1][:-1], dhq[1:], self.G(pd + 'attn.attn.weight'), self.G(pd + 'attn.attn.bias'), M=M1, K=H, N=H, ldw=2 * H)
        _wg(enc, deproj, self.G(pd + 'attn.attn.weight'), None, M=B * Tm, K=H, N=H, ldw=2 * H, dw_off=H)
        _dlin(deproj, Wa, denc, M=B * Tm, K=H, N=H, ldw=2 * H, w_off=H, accumulate=True)
        # ---- encoder backward through time
        dlo = ws.get('enc.dlo', (B * Tm, 2 * H))
        ops.dup_halves(denc, dlo, B * Tm, H)                                           # outputs = fwd + rev (:58)
        dgh_e = ws.get('enc.dgh', (Tm, B, 3 * H))
        dcar = ws.get('enc.dcarry', (B, H))
        for l in range(L - 1, -1, -1):
            I = E if l == 0 else 2 * H
            if l == 0:
                inp = ws['enc.emb']
            elif masks is not None and f'enc{l - 1}' in masks:
                inp = ws[f'enc.drop{l - 1}']
            else:
                inp = ws[f'enc.out{l - 1}']
            dinp = ws.get(f'enc.dinp{l}', (B * Tm, I))
            for d, suf in enumerate(('', '_reverse')):
                q = f'encoder.gru.%s_l{l}{suf}'
                hs, svd = ws[f'enc.hs{l}_{d}'], ws[f'enc.saved{l}_{d}']
                dgi_e = ws.get(f'enc.dgi{l}_{d}', (B * Tm, 3 * H))
                # carry into the last step: final hidden state 2*l + d initialises decoder layer 2*l + d if that is < n_layers (:241)
                if 2 * l + d < L:
                    dcar.copy_(dH[2 * l + d])
                else:
                    dcar.zero_()
                for step in range(Tm - 1, -1, -1):
                    t = step if d == 0 else Tm - 1 - step
                    ops.gru_gates_bwd(dcar, dlo[t:, d * H:], Tm * 2 * H, svd[step], B * H, hs[step], lengths, t, dgi_e[t:], Tm * 3 * H,
                                      dgh_e[step], dcar, B, H)
                    _dlin(dgh_e[step], self.P(q % 'weight_hh'), dcar, M=B, K=H, N=3 * H, accumulate=True)
                _wg(hs[:Tm], dgh_e, self.G(q % 'weight_hh'), self.G(q % 'bias_hh'), M=Tm * B, K=H, N=3 * H)
                _wg(inp, dgi_e, self.G(q % 'weight_ih'), self.G(q % 'bias_ih'), M=B * Tm, K=I, N=3 * H)
                _dlin(dgi_e, self.P(q % 'weight_ih'), dinp, M=B * Tm, K=I, N=3 * H, accumulate=(d == 1))
            if l > 0:
                if masks is not None and f'enc{l - 1}' in masks:
                    ops.mul(dinp, masks[f'enc{l - 1}'], dinp, B * Tm * 2 * H)
                dlo = dinp
            else:
                emb_p = self.arena.params['encoder.embedding.weight']
                if emb_p.requires_grad:
                    ops.embedding_scatter_add(dinp, ws['enc.idx'], None, self.G('encoder.embedding.weight'), B * Tm, E)

    def clip_and_step(self, optim, max_norm, host_step=True, world=1):
        """torch.nn.utils.clip_grad_norm_(parameters, max_norm) + optimizer.step() (train_seq2seq.py:48-49).  world > 1 (data parallel): the
        arena holds the SUM of the ranks' gradients; the norm of their mean is |sum| / world, so the sum is clipped against max_norm * world
        and the mean is taken inside Adam (grad_scale = 1 / world)."""
        ws = self.ws
        ss = ws.get('opt.sumsq', (1,), torch.float64)
        ss.zero_()
        ops.sumsq(self.arena.grad, self.arena.numel, ss)
        ops.clip_scale(self.arena.grad, self.arena.numel, ss, max_norm * world)
        self.arena.adam_step(optim, grad_scale=1.0 / world, host_step=host_step)
