"""Launch plans of the Speech2Gesture baseline (reference: scripts/model/speech2gesture.py:133-250, scripts/train_eval/train_speech2gesture.py).

Every Conv2d_tf / Conv1d_tf is  im2col (tg_im2col2d, TensorFlow SAME / VALID padding)  ->  GEMM on the column matrix (tcgen05 TF32 tiles in the
fast mode wherever TMA can describe the operands, fp32 FFMA otherwise) ; its backward is the weight-gradient GEMM on the SAME column matrix and
a column GEMM followed by tg_col2im2d.  BatchNorm (train-mode statistics, running-buffer updates) and LeakyReLU reuse the generator's
kernels; activations are channels-last [B,H,W,C] (a sequence is H = 1), so the reference's transposes do not exist.  The U-Net wiring
(bilinear make_1d, UnetUp's repeat_interleave + add, the skip connections' gradient fan-in), the pose differencing and the losses are the small
kernels of csrc/s2g.cu.  Parameters / gradients / Adam moments live in flat arenas bound to the caller's torch.optim.Adam."""
from __future__ import annotations

from typing import Optional

import torch

from . import config, ops
from .arena import ParamArena
from .engine import BN_EPS, BN_MOM, Workspace, mm_nn, mm_nt, wgrad

SLOPE = 0.2


def _same_pad(size, k, s):
    """TensorFlow SAME (speech2gesture.py:19-30): output ceil(size / s); the odd padding element goes to the far edge."""
    out = (size + s - 1) // s
    total = max(0, (out - 1) * s + k - size)
    return out, total // 2


class _ConvBlock:
    """One Conv{1,2}d_tf [+ BatchNorm + LeakyReLU(0.2)] of a plan: owns its column matrix, pre-activation output and activation."""

    def __init__(self, eng, tag, conv_name, bn_name, act_slope):
        self.e, self.tag, self.conv, self.bn, self.slope = eng, tag, conv_name, bn_name, act_slope

    def forward(self, x, B, H, W, training):
        e, ws = self.e, self.e.ws
        w = e.P(self.conv + '.weight')
        cout, cin = w.shape[0], w.shape[1]
        if w.dim() == 4:
            kh, kw = w.shape[2], w.shape[3]
        else:
            kh, kw = 1, w.shape[2]
        m = e.mod(self.conv)
        sh = m.stride[0] if w.dim() == 4 else 1
        sw = m.stride[-1]
        valid = getattr(m, 'padding', 0) == 'VALID' or getattr(m, 'padding', 0) in (0, (0,), (0, 0))
        if valid:
            Ho, Wo, pt, pl = (H - kh) // sh + 1, (W - kw) // sw + 1, 0, 0
        else:
            Ho, pt = _same_pad(H, kh, sh)
            Wo, pl = _same_pad(W, kw, sw)
        M, K = B * Ho * Wo, kh * kw * cin
        self.geom = (B, H, W, cin, kh, kw, sh, sw, pt, pl, Ho, Wo, cout)
        self.direct = kh == 1 and kw == 1 and sw == 1 and sh == 1
        if self.direct:
            col = x                                        # 1x1 convolution: the activation IS the column matrix
        else:
            col = ws.get(self.tag + '.col', (M, K))
            ops.im2col2d(x, col, B, H, W, cin, kh, kw, sh, sw, pt, pl, Ho, Wo)
        self.col = col
        w2 = ws.get(self.tag + '.w2', (cout, K)); w2t = ws.get(self.tag + '.w2t', (K, cout))
        ops.window_weights(w, w2, w2t, cout, cin, kh * kw)
        y = ws.get(self.tag + '.y', (M, cout))
        mm_nt(col, w2, y, M=M, N=cout, K=K, bias=e.P(self.conv + '.bias'))
        self.y = y
        out = y
        if self.bn is not None:
            sc, sh_ = e._bn_fwd(self.tag, y, M, cout, self.bn, training)
            out = ws.get(self.tag + '.a', (M, cout))
            ops.affine_lrelu(y, out, M, cout, sc, sh_, self.slope)
        elif self.slope is not None:
            out = ws.get(self.tag + '.a', (M, cout))
            ops.affine_lrelu(y, out, M, cout, e.ones(cout), e.zeros(cout), self.slope)
        return out, Ho, Wo, cout

    def backward(self, d, need_dx=True, param_grads=True):
        """d [M, cout] = gradient w.r.t. this block's output (overwritten).  Returns the gradient w.r.t. its input [B*H*W, cin] or None."""
        e, ws = self.e, self.e.ws
        B, H, W, cin, kh, kw, sh, sw, pt, pl, Ho, Wo, cout = self.geom
        M, K = B * Ho * Wo, kh * kw * cin
        if self.bn is not None:
            e._bn_bwd(self.tag, d, self.y, M, cout, self.bn, self.slope)
        elif self.slope is not None:
            ops.lrelu_bwd(d, self.y, d, M * cout, self.slope)
        if param_grads:
            dw2 = ws.get(self.tag + '.dw2', (cout, K)); dw2.zero_()
            wgrad(self.col, d, dw2, B=B, T=Ho * Wo, N=cout, Cin=K, dbias=e.G(self.conv + '.bias'))
            ops.window_wgrad_add(dw2, e.G(self.conv + '.weight'), cout, cin, kh * kw)
        if not need_dx:
            return None
        if self.direct:
            dx = ws.get(self.tag + '.dx', (M, K))
            mm_nn(d, ws[self.tag + '.w2'], ws[self.tag + '.w2t'] if config.fast() else None, dx, M=M, N=cout, K=K)
            return dx
        dcol = e.scratch(M * K).view(M, K)                 # one grow-only scratch for every layer's column gradient (up to 1.2 GB at batch 128)
        mm_nn(d, ws[self.tag + '.w2'], ws[self.tag + '.w2t'] if config.fast() else None, dcol, M=M, N=cout, K=K)
        dx = ws.get(self.tag + '.dx', (B * H * W, cin))
        ops.col2im2d(dcol, dx, B, H, W, cin, kh, kw, sh, sw, pt, pl, Ho, Wo)
        return dx



class _S2GBase:
    def __init__(self, module):
        self.m = module
        self.arena = ParamArena(module)
        self.ws: Optional[Workspace] = None
        self._mods = dict(module.named_modules())
        self._const = {}

    def ensure(self, device):
        self.arena.ensure(device)
        if self.ws is None or self.ws.device != device:
            self.ws = Workspace(device)
            self._const = {}
        self.bufs = dict(self.m.named_buffers())
        return self

    def P(self, name):
        return self.arena.params[name].data

    def G(self, name):
        return self.arena.gview(name)

    def mod(self, name):
        return self._mods[name]

    def scratch(self, n):
        buf = self._const.get('scratch')
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, device=self.ws.device)
            self._const['scratch'] = buf
        return buf[:n]

    def ones(self, n):
        k = ('1', n)
        if k not in self._const:
            self._const[k] = torch.ones(n, device=self.ws.device)
        return self._const[k]

    def zeros(self, n):
        k = ('0', n)
        if k not in self._const:
            self._const[k] = torch.zeros(n, device=self.ws.device)
        return self._const[k]

    def _bn_fwd(self, tag, y, M, C, bn, training):
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
        ws = self.ws
        sums = ws.get(tag + '.bsums', (2 * C,), torch.float64); sums.zero_()
        mean, rstd, scale, shift = ws[tag + '.mean'], ws[tag + '.rstd'], ws[tag + '.scale'], ws[tag + '.shift']
        ops.bn_bwd_reduce(d, y, M, C, mean, rstd, scale, shift, slope, sums)
        ops.bn_bwd_apply(d, y, d, M, C, mean, rstd, scale, shift, slope, self.P(bn + '.weight'), sums, self.G(bn + '.weight'), self.G(bn + '.bias'))

    def block(self, tag, seq_name, with_bn=True):
        """A ConvNormRelu nn.Sequential named `seq_name` (conv = .0, norm = .1)."""
        return _ConvBlock(self, tag, seq_name + '.0', seq_name + '.1' if with_bn else None, SLOPE)


class S2GGeneratorEngine(_S2GBase):
    """speech2gesture.Generator (speech2gesture.py:198-229) incl. its AudioEncoder U-Net (:133-195)."""

    def __init__(self, module):
        super().__init__(module)
        ae = 'audio_encoder.'
        self.first = [self.block('g.f%d' % i, ae + 'first_net.%d' % i) for i in range(8)]
        self.down1 = [self.block('g.d1_%d' % i, ae + 'down1.%d' % i) for i in range(2)]
        self.down = [self.block('g.d%d' % i, ae + 'down%d' % i) for i in range(2, 7)]
        self.up = [self.block('g.u%d' % i, ae + 'up%d.conv' % i) for i in range(1, 6)]
        self.dec = [self.block('g.dec%d' % i, 'decoder.%d' % i) for i in range(4)]
        self.final = _ConvBlock(self, 'g.final', 'final_out', None, None)

    def forward(self, in_spec, pre_poses, training):
        """in_spec [B, n_mel, L], pre_poses [B, n_pre, D] -> poses [B, n_poses, D] (a view into the workspace)."""
        ws = self.ws
        B, H, W = in_spec.shape
        T = self.m.gen_length
        x = in_spec.contiguous().float()
        self.ctx = dict(B=B, training=training)
        c = 1
        for blk in self.first:                                            # 8 x Conv2d_tf + BatchNorm2d + LeakyReLU
            x, H, W, c = blk.forward(x, B, H, W, training)
        self.ctx['HW8'] = (H, W)
        x1 = ws.get('g.x1', (B * T, c))                                    # make_1d: bilinear to (n_frames, 1)
        ops.resize_bilinear_fwd(x, x1, B, H, W, c, T, 1)
        h = x1
        for blk in self.down1:
            h, _, _, _ = blk.forward(h, B, 1, T, training)
        skips = [(h, T)]                                                   # x2 .. x6 (and x7 = the last element)
        t = T
        for blk in self.down:
            h, _, t, _ = blk.forward(h, B, 1, t, training)
            skips.append((h, t))
        self.skip_T = [s[1] for s in skips]
        xcur, tcur = skips[-1]
        for i, blk in enumerate(self.up):                                  # up1(x7, x6) ... up5(., x2)
            x2, t2 = skips[-2 - i]
            s = ws.get('g.us%d' % i, (B * t2, c))
            ops.upsample2_add_fwd(xcur, x2, s, B, tcur, t2, c)
            xcur, _, tcur, _ = blk.forward(s, B, 1, t2, training)
        # pre-pose encoder: Linear -> BatchNorm1d -> ReLU -> Linear
        npre = pre_poses.shape[1] * pre_poses.shape[2]
        pp = pre_poses.contiguous().float().view(B, npre)
        self.ctx['pp'] = pp
        p0 = ws.get('g.p0', (B, 32)); p0a = ws.get('g.p0a', (B, 32)); p1 = ws.get('g.p1', (B, 16))
        ops.linear(pp, self.P('pre_pose_encoder.0.weight'), self.P('pre_pose_encoder.0.bias'), p0, M=B, K=npre, N=32)
        sc, sh = self._bn_fwd('g.pbn', p0, B, 32, 'pre_pose_encoder.1', training)
        ops.affine_lrelu(p0, p0a, B, 32, sc, sh, 0.0)
        ops.linear(p0a, self.P('pre_pose_encoder.3.weight'), self.P('pre_pose_encoder.3.bias'), p1, M=B, K=32, N=16)
        feat = ws.get('g.feat', (B * T, c + 16))
        ops.concat_bcast_fwd(xcur, p1, feat, B, T, c, 16)
        h = feat
        for blk in self.dec:
            h, _, _, _ = blk.forward(h, B, 1, T, training)
        out, _, _, D = self.final.forward(h, B, 1, T, training)
        self.ctx.update(T=T, C=c, D=D)
        return out.view(B, T, D)

    def backward(self, d_out):
        """d_out [B, n_poses, D] (overwritten).  Accumulates every parameter gradient into the flat arena."""
        ws, c = self.ws, self.ctx
        B, T, C = c['B'], c['T'], c['C']
        d = self.final.backward(d_out.view(B * T, -1))
        for blk in reversed(self.dec):
            d = blk.backward(d)
        da = ws.get('g.dfa', (B * T, C)); dp1 = ws.get('g.dp1', (B, 16))
        ops.concat_bcast_bwd(d, da, dp1, B, T, C, 16)
        # pre-pose encoder
        npre = c['pp'].shape[1]
        dp0 = ws.get('g.dp0', (B, 32))
        ops.linear_wgrad(ws['g.p0a'], dp1, self.G('pre_pose_encoder.3.weight'), self.G('pre_pose_encoder.3.bias'), M=B, K=32, N=16)
        ops.linear_dgrad(dp1, self.P('pre_pose_encoder.3.weight'), dp0, M=B, K=32, N=16)
        self._bn_bwd('g.pbn', dp0, ws['g.p0'], B, 32, 'pre_pose_encoder.1', 0.0)
        ops.linear_wgrad(c['pp'], dp0, self.G('pre_pose_encoder.0.weight'), self.G('pre_pose_encoder.0.bias'), M=B, K=npre, N=32)
        # U-Net: up5 .. up1, collecting the skip gradients
        nsk = len(self.skip_T)
        dskip = [None] * nsk                                               # gradient flowing into skips[i] from the up path
        dcur = da
        for i in range(len(self.up) - 1, -1, -1):
            ds = self.up[i].backward(dcur)                                 # gradient w.r.t. the sum s = up(x1) + x2, [B*t2, C]
            j = nsk - 2 - i                                                # the skip this level added
            dskip[j] = ds
            t1 = self.skip_T[nsk - 1 - i]                                  # length of the level's upsampled input (x7 for up1)
            dx1 = ws.get('g.dup%d' % i, (B * t1, C))
            ops.upsample2_bwd(ds, dx1, B, t1, self.skip_T[j], C)
            dcur = dx1
        # dcur = gradient w.r.t. x7 (the deepest feature) from up1; walk down6 .. down2 adding each skip's gradient
        d = dcur
        for i in range(len(self.down) - 1, -1, -1):
            d = self.down[i].backward(d)                                   # gradient w.r.t. skips[i] from the down path
            ops.add(d, dskip[i], d, d.numel())
        for blk in reversed(self.down1):
            d = blk.backward(d)
        H8, W8 = c['HW8']
        dx8 = ws.get('g.dx8', (B * H8 * W8, C))
        ops.resize_bilinear_bwd(d, dx8, B, H8, W8, C, T, 1)
        d = dx8
        for i in range(len(self.first) - 1, -1, -1):
            d = self.first[i].backward(d, need_dx=i > 0)


class S2GDiscriminatorEngine(_S2GBase):
    """speech2gesture.Discriminator (speech2gesture.py:232-250): pose differences -> 4 Conv1d_tf."""

    def __init__(self, module):
        super().__init__(module)
        self.blocks = [_ConvBlock(self, 'sd.c0', 'net.0', None, SLOPE), self.block('sd.c1', 'net.2'), self.block('sd.c2', 'net.3'),
                       _ConvBlock(self, 'sd.c3', 'net.4', None, None)]

    def forward(self, poses, training, slot=''):
        """poses [B,T,D] -> scores [B*T', 1] (channels-last).  `slot` keeps the activations of concurrent passes apart."""
        ws = self.ws
        B, T, D = poses.shape
        for i, blk in enumerate(self.blocks):
            blk.tag = 'sd%s.c%d' % (slot, i)
        motion = ws.get('sd%s.motion' % slot, (B * (T - 1), D))
        ops.time_diff_fwd(poses, motion, B, T, D)
        h, t = motion, T - 1
        for blk in self.blocks:
            h, _, t, _ = blk.forward(h, B, 1, t, training)
        self.ctx = dict(B=B, T=T, D=D, slot=slot, geoms=[blk.geom for blk in self.blocks], cols=[blk.col for blk in self.blocks],
                        ys=[blk.y for blk in self.blocks])
        return h

    def backward(self, d_scores, need_dposes, param_grads=True, ctx=None, accumulate_into=None):
        """d_scores [B*T', 1] (overwritten).  Returns d poses [B,T,D] (added to `accumulate_into` when given) or None."""
        c = ctx if ctx is not None else self.ctx
        B, T, D, slot = c['B'], c['T'], c['D'], c['slot']
        for i, blk in enumerate(self.blocks):
            blk.tag = 'sd%s.c%d' % (slot, i)
            blk.geom, blk.col, blk.y = c['geoms'][i], c['cols'][i], c['ys'][i]
        d = d_scores
        for i in range(len(self.blocks) - 1, -1, -1):
            d = self.blocks[i].backward(d, need_dx=(i > 0 or need_dposes), param_grads=param_grads)
        if not need_dposes:
            return None
        if accumulate_into is not None:
            ops.time_diff_bwd(d, accumulate_into, B, T, D, accumulate=True)
            return accumulate_into
        dposes = self.ws.get('sd%s.dposes' % slot, (B, T, D))
        ops.time_diff_bwd(d, dposes, B, T, D)
        return dposes
