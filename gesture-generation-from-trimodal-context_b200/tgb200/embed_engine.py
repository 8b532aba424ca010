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

from . import ops
from .arena import ParamArena
from .engine import BN_EPS, BN_MOM, Workspace, _conv_out

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
    def backward(self, d_rec):
        """d_rec [B*34,27] = d loss / d recon.  Accumulates every parameter gradient into the (caller-zeroed) flat gradient arena."""
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
        self.encode_backward(dmu)
