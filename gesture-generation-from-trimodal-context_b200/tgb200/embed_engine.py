"""Launch plan for TRAINING the FGD auto-encoder: EmbeddingNet(mode='pose') in train mode (batch-statistics BatchNorm), the L1
reconstruction loss of train_feature_extractor.py:54-97 / train_joint_embed.py:5-51, the hand-derived backward and one flat Adam.

Reference anchors: scripts/model/embedding_net.py:42-82 (PoseEncoderConv), :165-217 (PoseDecoderConv), :276-308 (EmbeddingNet.forward).

Layout: activations are channels-last ([B,T,C] == row-major [B*T, C]).  The reference flattens / views channel-major tensors
twice (`out.flatten(1)` on [B,32,12], embedding_net.py:71; `out.view(B, 4, -1)`, :213); both become one tg_transpose_batched_f32 of
a [B, 12, 32] / [B, 4, 34] block.  BatchNorm + LeakyReLU outputs are never materialised: they are the operand prologue
(pscale / pshift / pslope) of the consuming GEMM, in the forward and in the weight gradient.  A ConvTranspose1d(k=3, stride 1) is the
data-gradient form of a convolution: forward = ops.conv1d_dgrad, its data gradient = ops.conv1d, its weight gradient = the
implicit-GEMM weight-gradient kernel with dilation -1 (rows x[s], dy[s+j]).  Everything is small (B=128: < 4 MB of activations,
190 k parameters), i.e. launch-latency bound: the whole step is ~75 launches and is replayed as one CUDA graph by the step function."""
from __future__ import annotations

from typing import Optional

import torch

from . import config, ops
from .arena import ParamArena
from .engine import BN_EPS, BN_MOM, S_WAVW, S_WGRAD, GeneratorEngine, GruPlan, Workspace, _conv_out, gru_arena_order, side

ENC_CONVS = ((3, 1), (3, 1), (4, 2))          # ConvNormRelu x3: (kernel, stride), embedding_net.py:20-26,46-48
ENC_SLOPE = DEC_SLOPE = 0.2                   # nn.LeakyReLU(0.2) :31,208,211 ; nn.LeakyReLU(True) == slope 1.0 == identity :57,60,203


class _PlanBase:
    """What every plan below needs: a workspace, the flat-arena accessors P (parameter) / G (gradient view) and the BatchNorm buffers."""
    arena: ParamArena
    ws: Optional[Workspace] = None

    def P(self, name):
        return self.arena.params[name].data

    def G(self, name):
        p = self.arena.params[name]
        return self.arena.gview(name) if p.requires_grad else None

    # ------------------------------------------------------------------------------------------------ BatchNorm helpers
    def _bn_fwd(self, tag, y, M, C, bn, training):
        """Batch statistics of y [M,C] (+ running-statistics update) -> per-channel (scale, shift) for the consumer's prologue."""
        ws = self.ws
        scale, shift = ws.get(tag + '.scale', (C,)), ws.get(tag + '.shift', (C,))
        if training:
            mean, rstd = ws.get(tag + '.mean', (C,)), ws.get(tag + '.rstd', (C,))
            sums = ws.get(tag + '.sums', (2 * C,), torch.float64); sums.zero_()
            ops.col_stats(y, C, M, C, sums)
            ops.bn_finalize(sums, M, C, BN_EPS, BN_MOM, 1, self.P(bn + '.weight'), self.P(bn + '.bias'), self.bufs[bn + '.running_mean'],
                            self.bufs[bn + '.running_var'], self.bufs[bn + '.num_batches_tracked'], mean, rstd, scale, shift)
        else:
            ops.bn_eval_fold(self.P(bn + '.weight'), self.P(bn + '.bias'), self.bufs[bn + '.running_mean'], self.bufs[bn + '.running_var'],
                             BN_EPS, None, scale, shift, C)
        return scale, shift

    def _bn_bwd(self, tag, d, y, M, C, bn, slope):
        """d [M,C] = gradient w.r.t. lrelu(bn(y)); overwritten with the gradient w.r.t. y; gamma / beta gradients accumulated."""
        ws = self.ws
        sums = ws.get(tag + '.bsums', (2 * C,), torch.float64); sums.zero_()
        mean, rstd, scale, shift = ws[tag + '.mean'], ws[tag + '.rstd'], ws[tag + '.scale'], ws[tag + '.shift']
        ops.bn_bwd_reduce(d, y, M, C, mean, rstd, scale, shift, slope, sums)
        ops.bn_bwd_apply(d, y, d, M, C, mean, rstd, scale, shift, slope, self.P(bn + '.weight'), sums, self.G(bn + '.weight'),
                         self.G(bn + '.bias'))



class PoseEncoderPlan(_PlanBase):
    """PoseEncoderConv (embedding_net.py:42-82) forward / backward; ENC = parameter-name prefix inside self.arena."""
    ENC = 'pose_encoder.'

    def encode(self, poses, training=True):
        """poses [B,34,27] -> (mu [B,32], logvar [B,32]); feature = mu (variational_encoding=False)."""
        ws = self.ws
        B, T, D = poses.shape
        e = self.ENC
        self.ctx = dict(B=B, T=T, D=D, poses=poses, training=training)
        x, tin, cin, pro = poses, T, D, {}
        self.enc_T = [T]
        for i, (k, s) in enumerate(ENC_CONVS):
            w = self.P(f'{e}net.{i}.0.weight')
            cout = w.shape[0]
            tout = _conv_out(tin, k, s)
            y = ws.get(f'ae.y{i}', (B * tout, cout))
            ops.conv1d(x, w, self.P(f'{e}net.{i}.0.bias'), y, B=B, Tin=tin, Cin=cin, N=cout, k=k, stride=s, **pro)
            sc, sh = self._bn_fwd(f'ae.y{i}', y, B * tout, cout, f'{e}net.{i}.1', training)
            pro = dict(pscale=sc, pshift=sh, pslope=ENC_SLOPE)
            x, tin, cin = y, tout, cout
            self.enc_T.append(tout)
        w = self.P(e + 'net.3.weight')
        cout, k = w.shape[0], w.shape[2]
        tout = _conv_out(tin, k, 1)
        y3 = ws.get('ae.y3', (B * tout, cout))
        ops.conv1d(x, w, self.P(e + 'net.3.bias'), y3, B=B, Tin=tin, Cin=cin, N=cout, k=k, **pro)
        self.enc_T.append(tout)
        nf = cout * tout
        w0 = self.P(e + 'out_net.0.weight')
        assert w0.shape[1] == nf, 'PoseEncoderConv.out_net is hard-wired to 34-frame clips (embedding_net.py:54-55)'
        f = ws.get('ae.f', (B, nf))
        ops.transpose_batched(y3, f, B, tout, cout)                              # [B,12,32] -> channel-major flatten [B,32*12]
        n0 = w0.shape[0]
        h0 = ws.get('ae.h0', (B, n0))
        ops.linear(f, w0, self.P(e + 'out_net.0.bias'), h0, M=B, K=nf, N=n0)
        sc, sh = self._bn_fwd('ae.h0', h0, B, n0, e + 'out_net.1', training)
        w1 = self.P(e + 'out_net.3.weight'); n1 = w1.shape[0]
        h1 = ws.get('ae.h1', (B, n1))
        ops.linear(h0, w1, self.P(e + 'out_net.3.bias'), h1, M=B, K=n0, N=n1, pscale=sc, pshift=sh, pslope=1.0)
        sc, sh = self._bn_fwd('ae.h1', h1, B, n1, e + 'out_net.4', training)
        w2 = self.P(e + 'out_net.6.weight'); n2 = w2.shape[0]
        h2 = ws.get('ae.h2', (B, n2))
        ops.linear(h1, w2, self.P(e + 'out_net.6.bias'), h2, M=B, K=n1, N=n2, pscale=sc, pshift=sh, pslope=1.0)
        mu, logvar = ws.get('ae.mu', (B, 32)), ws.get('ae.logvar', (B, 32))
        ops.linear(h2, self.P(e + 'fc_mu.weight'), self.P(e + 'fc_mu.bias'), mu, M=B, K=n2, N=32)
        ops.linear(h2, self.P(e + 'fc_logvar.weight'), self.P(e + 'fc_logvar.bias'), logvar, M=B, K=n2, N=32)
        return mu, logvar

    def encode_backward(self, dmu):
        """dmu [B,32] = d loss / d mu.  fc_logvar never reaches the loss when variational_encoding is False: its gradient stays zero."""
        ws, c = self.ws, self.ctx
        assert c['training'], 'backward through eval-mode BatchNorm is not on this path'
        B = c['B']
        e = self.ENC
        P, G = self.P, self.G
        # ---- encoder head: fc_mu, out_net (fc_logvar never reaches the loss: its gradient stays zero, train_feature_extractor.py:58-86)
        n2 = P(e + 'fc_mu.weight').shape[1]
        ops.linear_wgrad(ws['ae.h2'], dmu, G(e + 'fc_mu.weight'), G(e + 'fc_mu.bias'), M=B, K=n2, N=32)
        dh2 = ws.get('ae.d_h2', (B, n2))
        ops.linear_dgrad(dmu, P(e + 'fc_mu.weight'), dh2, M=B, K=n2, N=32)
        dprev = dh2
        for name, src, bn in (('out_net.6', 'ae.h1', 'out_net.4'), ('out_net.3', 'ae.h0', 'out_net.1')):
            w = P(e + name + '.weight'); n, kk = w.shape
            ops.linear_wgrad(ws[src], dprev, G(e + name + '.weight'), G(e + name + '.bias'), M=B, K=kk, N=n,
                             pscale=ws[src + '.scale'], pshift=ws[src + '.shift'], pslope=1.0)
            dsrc = ws.get(src.replace('ae.', 'ae.d_'), (B, kk))
            ops.linear_dgrad(dprev, w, dsrc, M=B, K=kk, N=n)
            self._bn_bwd(src, dsrc, ws[src], B, kk, e + bn, 1.0)
            dprev = dsrc
        w0 = P(e + 'out_net.0.weight'); n0, nf = w0.shape
        ops.linear_wgrad(ws['ae.f'], dprev, G(e + 'out_net.0.weight'), G(e + 'out_net.0.bias'), M=B, K=nf, N=n0)
        df = ws.get('ae.d_f', (B, nf))
        ops.linear_dgrad(dprev, w0, df, M=B, K=nf, N=n0)
        eT = self.enc_T                                                           # [34, 32, 30, 14, 12]
        c3 = P(e + 'net.3.weight').shape[0]
        dy = ws.get('ae.d_y3', (B * eT[4], c3))
        ops.transpose_batched(df, dy, B, c3, eT[4])                               # channel-major [B,32,12] -> channels-last [B,12,32]
        # ---- encoder convs, last first
        convs = [(f'{e}net.{i}.0', k, s) for i, (k, s) in enumerate(ENC_CONVS)] + [(e + 'net.3', 3, 1)]
        for li in (3, 2, 1, 0):
            name, k, s = convs[li]
            w = P(name + '.weight'); cout, cin, _ = w.shape
            tin, tout = eT[li], eT[li + 1]
            if li > 0:
                x = ws[f'ae.y{li - 1}']
                pro = dict(pscale=ws[f'ae.y{li - 1}.scale'], pshift=ws[f'ae.y{li - 1}.shift'], pslope=ENC_SLOPE)
            else:
                x, pro = c['poses'], {}
            ops.conv1d_wgrad(x, dy, G(name + '.weight'), G(name + '.bias'), B=B, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k, stride=s, **pro)
            if li == 0:
                break
            dx = ws.get(f'ae.d_y{li - 1}', (B * tin, cin))
            ops.conv1d_dgrad(dy, w, dx, B=B, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k, stride=s)
            self._bn_bwd(f'ae.y{li - 1}', dx, x, B * tin, cin, f'{e}net.{li - 1}.1', ENC_SLOPE)
            dy = dx


class AutoEncoderTrainEngine(PoseEncoderPlan):
    """EmbeddingNet(mode='pose'): PoseEncoderPlan + PoseDecoderConv (embedding_net.py:165-217) + loss."""

    def __init__(self, module):
        self.m = module
        self.arena = ParamArena(module)
        self.ws: Optional[Workspace] = None
        self.graph_slots = {}                 # captured CUDA graphs of the step (train_eval.train_joint_embed.ae_step)

    def ensure(self, device):
        if not self.arena.is_current():
            self.graph_slots = {}             # parameters were moved / re-created: captured pointers are stale
        self.arena.ensure(device)
        if self.ws is None or self.ws.device != device:
            self.ws = Workspace(device)
        self.bufs = dict(self.m.named_buffers())
        return self

    def forward(self, poses, training=True):
        """poses [B,34,27] -> (mu [B,32], logvar [B,32], recon [B,34,27])."""
        mu, logvar = self.encode(poses, training)
        return mu, logvar, self.decode(mu, training)

    def decode(self, mu, training=True):
        """PoseDecoderConv, length 34: mu [B,32] -> recon [B,34,27]."""
        ws = self.ws
        B, T, D = self.ctx['B'], self.ctx['T'], self.ctx['D']
        d = 'decoder.'
        wp = self.P(d + 'pre_net.0.weight'); c0 = wp.shape[0]
        g0 = ws.get('ae.g0', (B, c0))
        ops.linear(mu, wp, self.P(d + 'pre_net.0.bias'), g0, M=B, K=32, N=c0)
        sc, sh = self._bn_fwd('ae.g0', g0, B, c0, d + 'pre_net.1', training)
        wq = self.P(d + 'pre_net.3.weight'); c1 = wq.shape[0]
        g1 = ws.get('ae.g1', (B, c1))
        ops.linear(g0, wq, self.P(d + 'pre_net.3.bias'), g1, M=B, K=c0, N=c1, pscale=sc, pshift=sh, pslope=1.0)
        ch, L = 4, c1 // 4
        g1t = ws.get('ae.g1t', (B * L, ch))
        ops.transpose_batched(g1, g1t, B, ch, L)                                 # view [B,4,34] (channel-major) -> channels-last [B,34,4]
        x, tin, cin, pro = g1t, L, ch, {}
        self.dec_T = [L]
        for i, idx in enumerate((0, 3)):                                          # ConvTranspose1d(k=3) + BN + LeakyReLU(0.2), twice
            w = self.P(f'{d}net.{idx}.weight')                                    # [Cin, Cout, k]
            cout, k = w.shape[1], w.shape[2]
            tout = tin + k - 1
            y = ws.get(f'ae.t{i}', (B * tout, cout))
            ops.conv1d_dgrad(x, w, y, B=B, Tin=tout, Tout=tin, Cin=cout, N=cin, k=k, bias=self.P(f'{d}net.{idx}.bias'), **pro)
            sc, sh = self._bn_fwd(f'ae.t{i}', y, B * tout, cout, f'{d}net.{idx + 1}', training)
            pro = dict(pscale=sc, pshift=sh, pslope=DEC_SLOPE)
            x, tin, cin = y, tout, cout
            self.dec_T.append(tout)
        for i, idx in enumerate((6, 7)):
            w = self.P(f'{d}net.{idx}.weight')
            cout, k = w.shape[0], w.shape[2]
            tout = tin - k + 1
            y = ws.get(f'ae.c{i}', (B * tout, cout))
            ops.conv1d(x, w, self.P(f'{d}net.{idx}.bias'), y, B=B, Tin=tin, Cin=cin, N=cout, k=k, **pro)
            x, tin, cin, pro = y, tout, cout, {}
            self.dec_T.append(tout)
        assert tin == T and cin == D, (tin, cin)
        return x.view(B, T, D)

    # ------------------------------------------------------------------------------------------------ loss
    def loss(self, recon, target, use_diff, weight, acc, want_grad=True):
        """acc (fp64 [2], caller-zeroed) += (sum_b [mean|recon-target| (+ mean|frame differences|)], sum_b mean|recon-target|);
        returns d (weight*acc[0]) / d recon."""
        B, T, D = target.shape
        d_rec = self.ws.get('ae.d_c1', (B * T, D)) if want_grad else None
        ops.ae_recon_loss(recon, target, B, T, D, use_diff, weight, acc, d_rec)
        return d_rec

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, d_rec, d_mu_extra=None):
        """d_rec [B*34,27] = d loss / d recon (None: the loss does not read the reconstruction); d_mu_extra [B,32] = gradient reaching the
        latent from outside the decoder (module-level autograd API).  Accumulates every parameter gradient into the flat gradient arena."""
        if d_rec is None:
            if d_mu_extra is not None:
                self.encode_backward(d_mu_extra)
            return
        ws, c = self.ws, self.ctx
        assert c['training'], 'backward through eval-mode BatchNorm is not on this path'
        B = c['B']
        d = 'decoder.'
        P, G = self.P, self.G
        # ---- decoder.net.7 / net.6 (Conv1d k=3)
        dT = self.dec_T                                                           # [34, 36, 38, 36, 34]
        w = P(d + 'net.7.weight'); cout, cin, k = w.shape
        ops.conv1d_wgrad(ws['ae.c0'], d_rec, G(d + 'net.7.weight'), G(d + 'net.7.bias'), B=B, Tin=dT[3], Tout=dT[4], Cin=cin, N=cout, k=k)
        dc0 = ws.get('ae.d_c0', (B * dT[3], cin))
        ops.conv1d_dgrad(d_rec, w, dc0, B=B, Tin=dT[3], Tout=dT[4], Cin=cin, N=cout, k=k)
        w = P(d + 'net.6.weight'); cout, cin, k = w.shape
        pro = dict(pscale=ws['ae.t1.scale'], pshift=ws['ae.t1.shift'], pslope=DEC_SLOPE)
        ops.conv1d_wgrad(ws['ae.t1'], dc0, G(d + 'net.6.weight'), G(d + 'net.6.bias'), B=B, Tin=dT[2], Tout=dT[3], Cin=cin, N=cout, k=k, **pro)
        dy = ws.get('ae.d_t1', (B * dT[2], cin))
        ops.conv1d_dgrad(dc0, w, dy, B=B, Tin=dT[2], Tout=dT[3], Cin=cin, N=cout, k=k)
        # ---- the two ConvTranspose1d blocks, last first: x [B,tin,ci] -> y [B,tin+k-1,co]
        for i, idx in ((1, 3), (0, 0)):
            w = P(f'{d}net.{idx}.weight'); ci, co, k = w.shape
            tin, tout = dT[i], dT[i + 1]
            y = ws[f'ae.t{i}']
            self._bn_bwd(f'ae.t{i}', dy, y, B * tout, co, f'{d}net.{idx + 1}', DEC_SLOPE)       # dy: now d loss / d y
            if i == 1:
                x, pro = ws['ae.t0'], dict(pscale=ws['ae.t0.scale'], pshift=ws['ae.t0.shift'], pslope=DEC_SLOPE)
            else:
                x, pro = ws['ae.g1t'], {}
            # dW[ci,co,j] = sum_{b,s} x[b,s,ci] * dy[b,s+j,co]: rows of dy drive the loop, x is gathered at s = t - j (dilation -1)
            ops.conv_wgrad(x, dy, G(f'{d}net.{idx}.weight'), B=B, Tin=tin, Tout=tout, N=co, Cin=ci, taps=k, stride=1, dil=-1, pad=0,
                           ldw=k, wsj=1, wsc=co * k, dbias=G(f'{d}net.{idx}.bias'), **pro)
            dx = ws.get(f'ae.d_x{i}', (B * tin, ci))
            ops.conv1d(dy, w, None, dx, B=B, Tin=tout, Cin=co, N=ci, k=k)         # dx[s,ci] = sum_{j,co} dy[s+j,co] w[ci,co,j]
            dy = dx
        # ---- view [B,4,34] back to channel-major, decoder.pre_net
        c1 = P(d + 'pre_net.3.weight').shape[0]; c0 = P(d + 'pre_net.3.weight').shape[1]
        ch, L = 4, c1 // 4
        dg1 = ws.get('ae.d_g1', (B, c1))
        ops.transpose_batched(dy, dg1, B, L, ch)
        ops.linear_wgrad(ws['ae.g0'], dg1, G(d + 'pre_net.3.weight'), G(d + 'pre_net.3.bias'), M=B, K=c0, N=c1,
                         pscale=ws['ae.g0.scale'], pshift=ws['ae.g0.shift'], pslope=1.0)
        dg0 = ws.get('ae.d_g0', (B, c0))
        ops.linear_dgrad(dg1, P(d + 'pre_net.3.weight'), dg0, M=B, K=c0, N=c1)
        self._bn_bwd('ae.g0', dg0, ws['ae.g0'], B, c0, d + 'pre_net.1', 1.0)
        ops.linear_wgrad(ws['ae.mu'], dg0, G(d + 'pre_net.0.weight'), G(d + 'pre_net.0.bias'), M=B, K=32, N=c0)
        dmu = ws.get('ae.d_mu', (B, 32))
        ops.linear_dgrad(dg0, P(d + 'pre_net.0.weight'), dmu, M=B, K=32, N=c0)
        if d_mu_extra is not None:
            ops.add(dmu, d_mu_extra, dmu, dmu.numel())
        self.encode_backward(dmu)


# =====================================================================================================================
# Joint-embedding model: EmbeddingNet(mode != 'pose') = ContextEncoder + PoseEncoderConv + PoseDecoderGRU
# (scripts/model/embedding_net.py:130-162,220-308; trained by scripts/train_eval/train_joint_embed.py:5-51)
# =====================================================================================================================
class _Sub(_PlanBase):
    """P / G / BatchNorm helpers over the arena of ONE sub-module (parameter names relative to it)."""

    def __init__(self, arena, ws, bufs):
        self.arena, self.ws, self.bufs = arena, ws, bufs


class _PoseEncoderSub(PoseEncoderPlan):
    ENC = ''

    def __init__(self, arena, ws, bufs):
        self.arena, self.ws, self.bufs = arena, ws, bufs


class _ContextHost(_Sub):
    """ContextEncoder.text_encoder / .audio_encoder are the generator's TextEncoderTCN / WavEncoder (embedding_net.py:225-226) and carry
    the same parameter names relative to their parent, so the generator's launch plans for them (tgb200.engine.GeneratorEngine: causal
    two-tap tensor-core GEMMs, window-view strided convolutions, BatchNorm prologues, their backward) run on this arena unchanged."""
    WAV = GeneratorEngine.WAV
    wav_fast, wav_forward, wav_backward = GeneratorEngine.wav_fast, GeneratorEngine.wav_forward, GeneratorEngine.wav_backward
    text_forward, text_backward, make_masks = GeneratorEngine.text_forward, GeneratorEngine.text_backward, GeneratorEngine.make_masks
    _tcn_conv = staticmethod(GeneratorEngine._tcn_conv)
    _tcn_wgrad = staticmethod(GeneratorEngine._tcn_wgrad)
    _tcn_dgrad = staticmethod(GeneratorEngine._tcn_dgrad)

    def __init__(self, arena, ws, bufs, module):
        super().__init__(arena, ws, bufs)
        te = module.text_encoder
        self.use_text = True
        self.E = te.embedding.weight.shape[1]
        self.H = te.tcn.network[0].conv1.weight_v.shape[0]              # TCN channels (= args.hidden_size)
        self.n_tcn = len(te.tcn.network)
        self.tcn_k = te.tcn.network[0].conv1.weight_v.shape[2]
        self.p_emb, self.p_tcn = float(te.emb_dropout), float(te.tcn.network[0].dropout1.p)
        self.L, self.p_gru = 1, 0.0                                     # make_masks: text masks only

    def prep(self):
        """Per-optimiser-step derived weights: weight-normed TCN filters (tap-major, + per-tap transposes in fast mode)."""
        ws = self.ws
        for i in range(self.n_tcn):
            for j in (1, 2):
                q = f'text_encoder.tcn.network.{i}.conv{j}'
                v = self.P(q + '.weight_v')
                N, Cin, k = v.shape
                wT = ws.get(f'tcn.wT{i}_{j}', (k, Cin, N)) if config.fast() else None
                ops.weight_norm_fwd(v, self.P(q + '.weight_g'), ws.get(f'tcn.w{i}_{j}', (k, N, Cin)), wT, ws.get(f'tcn.inv{i}_{j}', (N,)), N, Cin, k)
        if config.fast():
            w = self.P('text_encoder.decoder.weight')
            ops.transpose(w, ws.get('T.text_encoder.decoder.weight', (w.shape[1], w.shape[0])), w.shape[0], w.shape[1])


class JointEmbeddingEngine:
    """Forward / hand-derived backward of the joint-embedding model.  Three flat arenas (context encoder, pose encoder, decoder): a step
    decodes ONE latent (embedding_net.py:295-303), only that branch and the decoder receive gradients, and torch.optim.Adam skips
    parameters without a gradient - moments and step counts included - so each arena carries its own Adam state and step counter.

    ContextEncoder.gru is unidirectional (H=256, 2 layers).  It runs on the bidirectional persistent recurrence kernels with the forward
    weights in both slots and an all-zero reverse half of gi / d_out: the reverse direction then computes values nobody reads and
    exactly-zero gradients; no new kernel, at the price of twice the (tiny) recurrence work."""
    HC, LC = 256, 2

    def __init__(self, module):
        self.m = module
        self.a_ctx = ParamArena(module.context_encoder, partial=True)
        self.a_pose = ParamArena(module.pose_encoder, partial=True)
        self.a_dec = ParamArena(module.decoder, gru_arena_order([n for n, _ in module.decoder.named_parameters()]), partial=True)
        self.ws: Optional[Workspace] = None
        self.noise = None                      # tests: dict(eps=[B,32], masks={...}, gru_masks=[...]) consumed by the next forward
        self.T = module.decoder.gen_length
        self.D = module.decoder.pose_dim

    def arenas(self):
        return self.a_ctx, self.a_pose, self.a_dec

    def ensure(self, device):
        rebuilt = not all(a.is_current() for a in self.arenas())
        for a in self.arenas():
            a.ensure(device)
        if self.ws is None or self.ws.device != device or rebuilt:
            self.ws = Workspace(device)
            m = self.m
            self.ctx = _ContextHost(self.a_ctx, self.ws, dict(m.context_encoder.named_buffers()), m.context_encoder)
            self.pose = _PoseEncoderSub(self.a_pose, self.ws, dict(m.pose_encoder.named_buffers()))
            self.dec = _Sub(self.a_dec, self.ws, dict(m.decoder.named_buffers()))
            self.gru = GruPlan(self.a_dec, 'gru', m.decoder.in_size, m.decoder.hidden_size, m.decoder.gru.num_layers, self.ws, 'jd')
        else:
            self.ctx.bufs = dict(self.m.context_encoder.named_buffers())
            self.pose.bufs = dict(self.m.pose_encoder.named_buffers())
            self.dec.bufs = dict(self.m.decoder.named_buffers())
        return self

    def _tc(self):
        """Recurrence kernel family of ContextEncoder.gru.  The strict-fp32 persistent kernels in BOTH arithmetic modes: this GRU is tiny
        (2 layers, H=256), so the tensor-core recurrence buys nothing measurable, and H=256 is a boundary of the tensor-core backward
        kernel's tiling (one vs two N halves) that has not been exercised on hardware, while the fp32 family has (forward and backward)."""
        return False

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, in_text, in_audio, pre_poses, poses, branch, training):
        """EmbeddingNet.forward (embedding_net.py:276-308), variational_encoding=False, branch in {'speech','pose'}."""
        noise, self.noise = self.noise, None
        ref = poses if poses is not None else pre_poses
        B = ref.shape[0]
        self.st = dict(B=B, branch=branch, training=training, noise=noise or {})
        r = dict(c_feat=None, c_mu=None, c_lv=None, p_mu=None, p_lv=None)
        if in_text is not None and in_audio is not None:
            r['c_feat'], r['c_mu'], r['c_lv'] = self._context_forward(in_text.contiguous(), in_audio.detach().contiguous().float(), B, training)
        if poses is not None:
            r['p_mu'], r['p_lv'] = self.pose.encode(poses.detach().contiguous().float(), training)
        latent = r['c_feat'] if branch == 'speech' else r['p_mu']
        assert latent is not None, 'the chosen branch has no input (embedding_net.py:298-301)'
        r['out'] = self._decoder_forward(latent, pre_poses, B, training)
        return r

    def _context_forward(self, in_text, in_audio, B, training):
        """ContextEncoder.forward (embedding_net.py:244-259) -> (z, mu, logvar)."""
        ws, h, T, HC = self.ws, self.ctx, self.T, self.HC
        M = B * T
        st = self.st
        st.update(in_text=in_text, in_audio=in_audio)
        h.prep()
        off = ws.get('jc.off', (1,), torch.int64, zero=True)
        masks = st['noise'].get('masks')
        if masks is None and training:
            masks = h.make_masks(B, T, self.m._noise_seed, off)
        st['masks'] = masks
        text = h.text_forward(in_text, B, T, masks)                                       # [M,32]
        audio = h.wav_forward(in_audio, training, 1)                                       # [M,32]
        x = ws.get('jc.x', (M, 64))
        ops.gru_input_concat(None, audio, text, None, x, B, B, T, 0, 32, 32, 0)            # cat((audio, text), dim=2), :250
        gi = ws.get('jc.gi', (M, 6 * HC), zero=True)                                       # reverse half is never written: stays zero
        tc = self._tc()
        sync = ws.get('jc.sync', (max(ops.gru_tf32_sync_ints(B, HC) if tc else ops.gru_sync_ints(B, HC), 1),), torch.int32)
        inp, K, lda = x, 64, 64
        for l in range(self.LC):
            whh, bhh = h.P(f'gru.weight_hh_l{l}'), h.P(f'gru.bias_hh_l{l}')
            whhT = ws.get(f'jc.whhT{l}', (HC, 3 * HC))
            ops.transpose(whh, whhT, 3 * HC, HC)
            ops.conv_gemm(inp, h.P(f'gru.weight_ih_l{l}'), gi, B=1, Tin=M, Tout=M, N=3 * HC, Cin=K, taps=1, lda=lda, ldw=K, wsc=1,
                          ldc=6 * HC, bias=h.P(f'gru.bias_ih_l{l}'))
            out = ws.get(f'jc.out{l}', (M, 2 * HC))
            saved = ws.get(f'jc.saved{l}', (4, M, 2 * HC)) if training else None
            if tc:
                ops.gru_layer_fwd_tf32(gi, whh, whh, bhh, bhh, out, saved, M * 2 * HC, sync, B, T, HC)
            else:
                ops.gru_layer_fwd(gi, whhT, whhT, bhh, bhh, out, saved, M * 2 * HC, sync, B, T, HC)
            inp, K, lda = out, HC, 2 * HC
        last = out.view(B, T, 2 * HC)[:, T - 1, :HC]                                       # output[:, -1] (:253): a strided view, pitch T*2H
        h0 = ws.get('jc.h0', (B, 128))
        ops.linear(last, h.P('out.0.weight'), h.P('out.0.bias'), h0, M=B, K=HC, N=128, lda=T * 2 * HC)
        sc, sh = h._bn_fwd('jc.h0', h0, B, 128, 'out.1', training)
        h1 = ws.get('jc.h1', (B, 32))
        ops.linear(h0, h.P('out.3.weight'), h.P('out.3.bias'), h1, M=B, K=128, N=32, pscale=sc, pshift=sh, pslope=0.0)     # BN + ReLU prologue
        mu, lv, z = ws.get('jc.mu', (B, 32)), ws.get('jc.lv', (B, 32)), ws.get('jc.z', (B, 32))
        ops.linear(h1, h.P('fc_mu.weight'), h.P('fc_mu.bias'), mu, M=B, K=32, N=32)
        ops.linear(h1, h.P('fc_logvar.weight'), h.P('fc_logvar.bias'), lv, M=B, K=32, N=32)
        eps = ws.get('jc.eps', (B, 32))
        if st['noise'].get('eps') is not None:
            eps.copy_(st['noise']['eps'])
        else:
            ops.philox_normal(eps, B * 32, self.m._noise_seed, off, 900)
        ops.increment_i64(off, 1)
        ops.reparam_fwd(mu, lv, eps, z, B * 32)                                            # unconditional, embedding_net.py:258
        return z, mu, lv

    def _decoder_forward(self, latent, pre_poses, B, training):
        """PoseDecoderGRU.forward (embedding_net.py:151-162): latent [B,32], pre_poses [B,4,D] -> [B,T,D]."""
        ws, d, T, D = self.ws, self.dec, self.T, self.D
        M = B * T
        H = self.gru.H
        pp = ws.get('jd.pp', (B, pre_poses.shape[1] * D))
        pp.view(B, -1, D).copy_(pre_poses.detach())
        Kp = pp.shape[1]
        p0 = ws.get('jd.p0', (B, 32))
        ops.linear(pp, d.P('pre_pose_net.0.weight'), d.P('pre_pose_net.0.bias'), p0, M=B, K=Kp, N=32)
        sc, sh = d._bn_fwd('jd.p0', p0, B, 32, 'pre_pose_net.1', training)
        feat = ws.get('jd.feat', (B, 64))                                                  # cat((pre_pose_feat, latent), dim=1), :153
        ops.linear(p0, d.P('pre_pose_net.3.weight'), d.P('pre_pose_net.3.bias'), feat, M=B, K=32, N=32, ldc=64, pscale=sc, pshift=sh, pslope=0.0)
        feat[:, 32:].copy_(latent)
        gin = ws.get('jd.gin', (M, 64))
        ops.gru_input_concat(None, None, None, feat, gin, B, 1, T, 0, 0, 0, 64)            # unsqueeze(1).repeat(1, T, 1), :154
        self.gru.prep()
        gm = self.st['noise'].get('gru_masks')
        p_drop = float(self.m.decoder.gru.dropout)
        if gm is None and training and p_drop > 0:
            off = ws.get('jd.off', (1,), torch.int64, zero=True)
            gm = []
            for l in range(self.gru.L - 1):
                mk = ws.get(f'jd.mask{l}', (M, 2 * H))
                ops.philox_dropout_mask(mk, M * 2 * H, p_drop, self.m._noise_seed, off, 950 + l)
                gm.append(mk)
            gm.append(None)
            ops.increment_i64(off, 1)
        self.st['gru_masks'] = gm
        out = self.gru.forward(gin, B, T, gm, save=training)                               # [M, 2H]
        hs, o1, rec = ws.get('jd.hs', (M, H)), ws.get('jd.o1', (M, H // 2)), ws.get('jd.rec', (M, D))
        ops.sum_halves(out, hs, M, H)                                                      # :157
        ops.linear(hs, d.P('out.0.weight'), d.P('out.0.bias'), o1, M=M, K=H, N=H // 2)     # LeakyReLU(True) == identity, :147
        ops.linear(o1, d.P('out.2.weight'), d.P('out.2.bias'), rec, M=M, K=H // 2, N=D)
        return rec.view(B, T, D)

    # ------------------------------------------------------------------------------------------------ loss / backward
    def loss(self, recon, target, acc, want_grad=True):
        B, T, D = target.shape
        d_rec = self.ws.get('jd.d_rec', (B * T, D)) if want_grad else None
        ops.ae_recon_loss(recon, target, B, T, D, False, 1.0, acc, d_rec)                  # train_joint_embed.py:21-29 (no frame-difference term)
        return d_rec

    def backward(self, d_rec):
        """Accumulates the gradients of the decoder and of the branch that produced the latent into their (caller-zeroed) arenas."""
        ws, d, st, T, D = self.ws, self.dec, self.st, self.T, self.D
        assert st['training']
        B = st['B']
        M = B * T
        H = self.gru.H
        do1, dhs, dout = ws.get('jd.do1', (M, H // 2)), ws.get('jd.dhs', (M, H)), ws.get('jd.dout', (M, 2 * H))
        ops.linear_wgrad(ws['jd.o1'], d_rec, d.G('out.2.weight'), d.G('out.2.bias'), M=M, K=H // 2, N=D)
        ops.linear_dgrad(d_rec, d.P('out.2.weight'), do1, M=M, K=H // 2, N=D)
        ops.linear_wgrad(ws['jd.hs'], do1, d.G('out.0.weight'), d.G('out.0.bias'), M=M, K=H, N=H // 2)
        ops.linear_dgrad(do1, d.P('out.0.weight'), dhs, M=M, K=H, N=H // 2)
        ops.dup_halves(dhs, dout, M, H)
        dgin = self.gru.backward(dout, ws['jd.gin'], B, 0, B, T, st['gru_masks'], True)    # [M,64]
        dfeat = ws.get('jd.dfeat', (B, 64))
        ops.gru_input_split_bwd(dgin, None, None, dfeat, B, T, 0, 0, 0, 64)                # repeat over T -> sum over T
        pro = dict(pscale=ws['jd.p0.scale'], pshift=ws['jd.p0.shift'], pslope=0.0)
        ops.linear_wgrad(ws['jd.p0'], dfeat, d.G('pre_pose_net.3.weight'), d.G('pre_pose_net.3.bias'), M=B, K=32, N=32, ldg=64, **pro)
        dp0 = ws.get('jd.dp0', (B, 32))
        ops.linear_dgrad(dfeat, d.P('pre_pose_net.3.weight'), dp0, M=B, K=32, N=32, lda=64)
        d._bn_bwd('jd.p0', dp0, ws['jd.p0'], B, 32, 'pre_pose_net.1', 0.0)
        pp = ws['jd.pp']
        ops.linear_wgrad(pp, dp0, d.G('pre_pose_net.0.weight'), d.G('pre_pose_net.0.bias'), M=B, K=pp.shape[1], N=32)
        dlat = ws.get('jd.dlat', (B, 32))
        dlat.copy_(dfeat[:, 32:])
        if st['branch'] == 'pose':
            self.pose.encode_backward(dlat)
        else:
            self._context_backward(dlat)
        side.join(S_WGRAD); side.join(S_WAVW)

    def _context_backward(self, dz):
        ws, h, st, T, HC = self.ws, self.ctx, self.st, self.T, self.HC
        B = st['B']
        M = B * T
        dmu, dlv = ws.get('jc.dmu', (B, 32)), ws.get('jc.dlv', (B, 32))
        dmu.zero_(); dlv.zero_()
        ops.reparam_bwd(dz, ws['jc.lv'], ws['jc.eps'], dmu, dlv, B * 32)
        h1, h0 = ws['jc.h1'], ws['jc.h0']
        ops.linear_wgrad(h1, dmu, h.G('fc_mu.weight'), h.G('fc_mu.bias'), M=B, K=32, N=32)
        ops.linear_wgrad(h1, dlv, h.G('fc_logvar.weight'), h.G('fc_logvar.bias'), M=B, K=32, N=32)
        dh1 = ws.get('jc.dh1', (B, 32))
        ops.linear_dgrad(dmu, h.P('fc_mu.weight'), dh1, M=B, K=32, N=32)
        ops.linear_dgrad(dlv, h.P('fc_logvar.weight'), dh1, M=B, K=32, N=32, accumulate=True)
        pro = dict(pscale=ws['jc.h0.scale'], pshift=ws['jc.h0.shift'], pslope=0.0)
        ops.linear_wgrad(h0, dh1, h.G('out.3.weight'), h.G('out.3.bias'), M=B, K=128, N=32, **pro)
        dh0 = ws.get('jc.dh0', (B, 128))
        ops.linear_dgrad(dh1, h.P('out.3.weight'), dh0, M=B, K=128, N=32)
        h._bn_bwd('jc.h0', dh0, h0, B, 128, 'out.1', 0.0)
        pitch = T * 2 * HC
        out1 = ws[f'jc.out{self.LC - 1}']
        last = out1.view(B, T, 2 * HC)[:, T - 1, :HC]
        ops.linear_wgrad(last, dh0, h.G('out.0.weight'), h.G('out.0.bias'), M=B, K=HC, N=128, lda=pitch)
        dout = ws.get('jc.dout1', (M, 2 * HC))
        dout.zero_()                                                                       # only output[:, -1] of the forward half is read
        ops.linear_dgrad(dh0, h.P('out.0.weight'), dout.view(B, T, 2 * HC)[:, T - 1, :HC], M=B, K=HC, N=128, ldc=pitch)
        tc = self._tc()
        partial = ws.get('jc.partial', (max(ops.gru_bwd_tf32_scratch_floats(B, HC) if tc else ops.gru_bwd_scratch_floats(B, HC), 1),))
        sync = ws.get('jc.bsync', (max(ops.gru_tf32_sync_ints(B, HC) if tc else ops.gru_sync_ints(B, HC), 1),), torch.int32)
        x = ws['jc.x']
        dx = None
        for l in range(self.LC - 1, -1, -1):
            whh = h.P(f'gru.weight_hh_l{l}')
            out, saved = ws[f'jc.out{l}'], ws[f'jc.saved{l}']
            dgi, dgh = ws.get(f'jc.dgi{l}', (M, 6 * HC)), ws.get(f'jc.dgh{l}', (M, 6 * HC))
            if tc:
                whhT = ws[f'jc.whhT{l}']
                ops.gru_layer_bwd_tf32(dout, out, saved[0], M * 2 * HC, whhT, whhT, dgi, dgh, partial, sync, B, T, HC)
            else:
                ops.gru_layer_bwd(dout, out, saved[0], M * 2 * HC, whh, whh, dgi, dgh, partial, sync, B, T, HC)
            inp, K, lda = (x, 64, 64) if l == 0 else (ws[f'jc.out{l - 1}'], HC, 2 * HC)
            ops.conv_wgrad(inp, dgi, h.G(f'gru.weight_ih_l{l}'), B=1, Tin=M, Tout=M, N=3 * HC, Cin=K, taps=1, lda=lda, ldg=6 * HC, ldw=K, wsc=1,
                           dbias=h.G(f'gru.bias_ih_l{l}'))
            # dW_hh += dgh^T h_{t-1}: the layer output one step earlier, zero at the start of a clip (pad = +1)
            ops.conv_wgrad(out, dgh, h.G(f'gru.weight_hh_l{l}'), B=B, Tin=T, Tout=T, N=3 * HC, Cin=HC, taps=1, pad=1, lda=2 * HC, ldg=6 * HC,
                           ldw=HC, wsc=1, dbias=h.G(f'gru.bias_hh_l{l}'))
            if l > 0:
                dprev = ws.get('jc.dout0', (M, 2 * HC), zero=True)                          # reverse half stays zero
                ops.linear_dgrad(dgi, h.P(f'gru.weight_ih_l{l}'), dprev, M=M, K=HC, N=3 * HC, lda=6 * HC, ldc=2 * HC)
                dout = dprev
            else:
                dx = ws.get('jc.dx', (M, 64))
                ops.linear_dgrad(dgi, h.P('gru.weight_ih_l0'), dx, M=M, K=64, N=3 * HC, lda=6 * HC)
        daud, dtxt = ws.get('jc.daud', (M, 32)), ws.get('jc.dtxt', (M, 32))
        ops.gru_input_split_bwd(dx, daud, dtxt, None, B, T, 0, 32, 32, 0)
        h.text_backward(dtxt, st['in_text'], 0, B, T, st['masks'])
        h.wav_backward(daud, st['in_audio'])
