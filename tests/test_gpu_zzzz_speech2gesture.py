"""B200: the Speech2Gesture baseline (reference scripts/model/speech2gesture.py + train_eval/train_speech2gesture.py)
through the C ABI: the plumbing kernels of csrc/s2g.cu against torch, the modules and two training steps against the reference-executed
golden (fp32 mode), and a batch-32 step against the fp64 oracle in both arithmetic modes."""
import argparse
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_l2
from oracle import s2g_oracle as SO
from oracle import synth
from oracle.make_golden import digest
from oracle.make_golden_s2g import B, D, D_LR_W, D_SEED, G_SEED, LR, N_PRE, T, W_GAN, W_REG, make_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _build(dev):
    from model.speech2gesture import Discriminator, Generator
    G, Dn = Generator(T, D, N_PRE), Discriminator(D)
    gsd, dsd = synth.s2g_state_dict(G.state_dict(), G_SEED), synth.s2g_state_dict(Dn.state_dict(), D_SEED)
    G.load_state_dict(gsd, strict=True); Dn.load_state_dict(dsd, strict=True)
    return G.to(dev), Dn.to(dev), gsd, dsd


def _close(a, ref, tol, what):
    a, ref = np.asarray(a), np.asarray(ref)
    scale = max(abs(ref[0]), 1e-9)
    assert abs(a[0] - ref[0]) <= tol * scale + 1e-7, (what, 'l2', a[0], ref[0])
    assert np.abs(a[2:] - ref[2:]).max() <= tol * max(np.abs(ref[2:]).max(), 1e-6) + 1e-6 * scale, what


@pytest.mark.parametrize('B_,H,W,C,kh,kw,sh,sw,same', [(3, 16, 9, 8, 3, 3, 1, 1, True), (2, 17, 11, 4, 4, 4, 2, 2, True), (2, 1, 34, 12, 1, 4, 1, 2, True),
                                                       (2, 14, 7, 8, 3, 3, 1, 1, False), (3, 1, 9, 5, 1, 3, 1, 1, True)])
def test_im2col_col2im_vs_torch_conv(dev, B_, H, W, C, kh, kw, sh, sw, same):
    """im2col -> matmul reproduces Conv2d_tf (SAME / VALID); col2im is its adjoint (checked against autograd)."""
    from tgb200 import ops
    from tgb200.s2g_engine import _same_pad
    g = torch.Generator().manual_seed(H * W + C)
    x = torch.randn(B_, H, W, C, generator=g).to(dev)
    N = 6
    w = torch.randn(N, C, kh, kw, generator=g).to(dev)
    if same:
        Ho, pt = _same_pad(H, kh, sh); Wo, pl = _same_pad(W, kw, sw)
    else:
        Ho, Wo, pt, pl = (H - kh) // sh + 1, (W - kw) // sw + 1, 0, 0
    col = torch.full((B_ * Ho * Wo, kh * kw * C), float('nan'), device=dev)
    ops.im2col2d(x, col, B_, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo)
    w2 = w.permute(0, 2, 3, 1).reshape(N, -1)
    y = (col.double() @ w2.double().t()).view(B_, Ho, Wo, N)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    ref = SO.conv_tf(xr, w.double(), None, (sh, sw), 'SAME' if same else 'VALID')
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < 1e-6
    dy = torch.randn(ref.shape, generator=g).to(dev).double()
    ref.backward(dy)
    dcol = (dy.permute(0, 2, 3, 1).reshape(-1, N) @ w2.double()).float().contiguous()
    dx = torch.full((B_, H, W, C), float('nan'), device=dev)
    ops.col2im2d(dcol, dx, B_, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo)
    assert rel_l2(dx.permute(0, 3, 1, 2), xr.grad) < 1e-5


def test_unet_plumbing_kernels_vs_torch(dev):
    from tgb200 import ops
    g = torch.Generator().manual_seed(3)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    # bilinear make_1d and its adjoint
    B_, H, W, C, Ho = 3, 14, 7, 8, 34
    x = r(B_, H, W, C)
    y = torch.empty(B_, Ho, 1, C, device=dev)
    ops.resize_bilinear_fwd(x, y, B_, H, W, C, Ho, 1)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(xr, size=(Ho, 1), mode='bilinear', align_corners=False)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < 1e-6
    dy = r(B_, Ho, 1, C)
    ref.backward(dy.double().permute(0, 3, 1, 2))
    dx = torch.full((B_, H, W, C), float('nan'), device=dev)
    ops.resize_bilinear_bwd(dy, dx, B_, H, W, C, Ho, 1)
    assert rel_l2(dx.permute(0, 3, 1, 2), xr.grad) < 1e-5
    # UnetUp: repeat_interleave + crop + add
    for T1, T2 in ((2, 3), (3, 5), (17, 34)):
        x1, x2 = r(B_, T1, C), r(B_, T2, C)
        out = torch.empty(B_, T2, C, device=dev)
        ops.upsample2_add_fwd(x1, x2, out, B_, T1, T2, C)
        assert torch.equal(out, torch.repeat_interleave(x1, 2, dim=1)[:, :T2] + x2)
        x1r = x1.double().requires_grad_(True)
        d = r(B_, T2, C)
        (torch.repeat_interleave(x1r, 2, dim=1)[:, :T2] * d.double()).sum().backward()
        dx1 = torch.empty(B_, T1, C, device=dev)
        ops.upsample2_bwd(d, dx1, B_, T1, T2, C)
        assert rel_l2(dx1, x1r.grad) < 1e-6
    # time differences
    p = r(B_, 9, 5)
    m = torch.empty(B_, 8, 5, device=dev)
    ops.time_diff_fwd(p, m, B_, 9, 5)
    assert torch.equal(m, p[:, 1:] - p[:, :-1])
    pr = p.double().requires_grad_(True)
    dm = r(B_, 8, 5)
    ((pr[:, 1:] - pr[:, :-1]) * dm.double()).sum().backward()
    dp = torch.ones(B_, 9, 5, device=dev)
    ops.time_diff_bwd(dm, dp, B_, 9, 5, accumulate=True)
    assert rel_l2(dp - 1.0, pr.grad) < 1e-5
    # concat with a broadcast feature
    a, q = r(B_, 6, 8), r(B_, 4)
    f = torch.empty(B_, 6, 12, device=dev)
    ops.concat_bcast_fwd(a, q, f, B_, 6, 8, 4)
    assert torch.equal(f, torch.cat((a, q.unsqueeze(1).expand(B_, 6, 4)), dim=2))
    df = r(B_, 6, 12)
    da, dq = torch.empty_like(a), torch.empty_like(q)
    ops.concat_bcast_bwd(df, da, dq, B_, 6, 8, 4)
    assert torch.equal(da, df[:, :, :8]) and rel_l2(dq, df[:, :, 8:].sum(1)) < 1e-6
    # losses
    s = r(40)
    sc = torch.zeros(2, dtype=torch.float64, device=dev)
    ds = torch.empty(40, device=dev)
    ops.mse_const(s, 40, 1.0, 10.0, sc[0:], ds)
    sr = s.double().requires_grad_(True)
    l = F.mse_loss(torch.ones_like(sr), sr); (10.0 * l).backward()
    assert abs(sc[0].item() - l.item()) < 1e-6 and rel_l2(ds, sr.grad) < 1e-6
    u, v = r(3, 7, 5), r(3, 7, 5)
    du = torch.empty_like(u)
    ops.l1_loss(u, v, u.numel(), 100.0, sc[1:], du)
    ur = u.double().requires_grad_(True)
    l = (ur - v.double()).abs().mean(); (100.0 * l).backward()
    assert abs(sc[1].item() - l.item()) < 1e-6 and rel_l2(du, ur.grad) < 1e-6


def test_eval_forward_vs_reference_golden(dev):
    from tgb200 import config
    g = np.load(os.path.join(GOLDEN, 's2g_step.npz'))
    old = config.set_mode('fp32')
    try:
        G, Dn, _, _ = _build(dev)
        G.eval(); Dn.eval()
        spec, target = make_inputs(B, 7)
        out = G(spec.to(dev), target[:, :N_PRE].to(dev))
        assert out.shape == (B, T, D)
        assert rel_l2(out, torch.from_numpy(g['eval/out'])) < 1e-4, rel_l2(out, torch.from_numpy(g['eval/out']))
        dis = Dn(target.to(dev))
        assert dis.shape == tuple(g['eval/dis'].shape)
        assert rel_l2(dis, torch.from_numpy(g['eval/dis'])) < 1e-4
    finally:
        config.set_mode(old)


def test_train_iter_vs_reference_golden(dev):
    """Two consecutive train_iter_speech2gesture calls (fp32 mode) vs the reference's own run: losses, every generator gradient, post-Adam
    weights and BatchNorm buffers of both networks."""
    from tgb200 import config
    from train_eval.train_speech2gesture import train_iter_speech2gesture
    g = np.load(os.path.join(GOLDEN, 's2g_step.npz'))
    old = config.set_mode('fp32')
    try:
        G, Dn, _, _ = _build(dev)
        G.train(); Dn.train()
        args = argparse.Namespace(n_pre_poses=N_PRE, loss_regression_weight=W_REG, loss_gan_weight=W_GAN)
        g_opt = torch.optim.Adam(G.parameters(), lr=LR, betas=(0.5, 0.999))
        d_opt = torch.optim.Adam(Dn.parameters(), lr=LR * D_LR_W, betas=(0.5, 0.999))
        for step in (1, 2):
            spec, target = torch.from_numpy(g[f's{step}/spec']).to(dev), torch.from_numpy(g[f's{step}/target']).to(dev)
            ret = train_iter_speech2gesture(args, spec, target, G, Dn, g_opt, d_opt, torch.nn.L1Loss())
            tol = 1e-4 if step == 1 else 2e-2      # step 2 starts from weights that went through two sign-like Adam updates (see below)
            for k, v in ret.items():
                r = float(g[f's{step}/loss/{k}'])
                assert abs(v - r) <= tol * abs(r) + 1e-7, (step, k, v, r)
            if step == 1:
                # The generator's gradient passes through the discriminator AFTER its first Adam step, which is sign-like (every weight moves
                # by +-lr): entries whose gradient is at round-off level move in different directions in any two fp32 implementations
                # (measured on the B200: D' differs from the fp64 oracle's by 2 lr in a few entries of net.0 / net.3, the GAN part of the
                # generator gradient by ~1e-2, while d_out agrees to 1e-6 with fp64 autograd through the SAME D').  Tight gradient parity is
                # therefore checked with the discriminator frozen (test_train_iter_batch32_vs_fp64_oracle); here: loosely, per tensor.
                for k, p in G.named_parameters():
                    ref = g[f's{step}/ggrad/{k}']
                    if ref[0] < 1e-4:
                        assert digest(p.grad.cpu())[0] < 1e-3, k
                        continue
                    _close(digest(p.grad.cpu()), ref, 5e-2, (step, 'ggrad', k))
                for k, v in G.state_dict().items():
                    if 'running' in k:
                        _close(digest(v.cpu()), g[f's{step}/gpost/{k}'], 1e-3, (step, 'gpost', k))
                for k, v in Dn.state_dict().items():
                    if 'running' in k:
                        _close(digest(v.cpu()), g[f's{step}/dpost/{k}'], 1e-3, (step, 'dpost', k))
            assert int(G.state_dict()['decoder.0.1.num_batches_tracked']) == step
            assert int(Dn.state_dict()['net.2.1.num_batches_tracked']) == 3 * step
    finally:
        config.set_mode(old)


@pytest.mark.parametrize('mode,tol', [('fp32', 2e-4), ('tf32', 1e-2)])
def test_train_iter_batch32_vs_fp64_oracle(dev, mode, tol):
    """One step at batch 32 against the fp64 oracle with the discriminator's learning rate set to 0 (its Adam step becomes the identity, so
    the generator step sees the same D on both sides - see the note in test_train_iter_vs_reference_golden): losses, generated poses, the
    discriminator step's weight gradients and every generator gradient."""
    D_LR_W = 0.0
    from tgb200 import config
    from train_eval.train_speech2gesture import train_iter_speech2gesture
    old = config.set_mode(mode)
    try:
        G, Dn, gsd, dsd = _build(dev)
        G.train(); Dn.train()
        Bb = 32
        spec, target = make_inputs(Bb, 41)
        f64 = lambda sd: {k: (v.to(dev).double() if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
        want = SO.train_iter_oracle(f64(gsd), f64(dsd), {}, {}, 1, spec.to(dev).double(), target.to(dev).double(), N_PRE, W_REG, W_GAN, LR, LR * D_LR_W)
        args = argparse.Namespace(n_pre_poses=N_PRE, loss_regression_weight=W_REG, loss_gan_weight=W_GAN)
        g_opt = torch.optim.Adam(G.parameters(), lr=LR, betas=(0.5, 0.999))
        d_opt = torch.optim.Adam(Dn.parameters(), lr=LR * D_LR_W, betas=(0.5, 0.999))
        ret = train_iter_speech2gesture(args, spec.to(dev), target.to(dev), G, Dn, g_opt, d_opt, None)
        for k, v in ret.items():
            assert abs(v - want['losses'][k]) <= tol * abs(want['losses'][k]) + 1e-7, (mode, k, v, want['losses'][k])
        out = G.engine().ws['g.final.y'].view(Bb, T, D)
        assert rel_l2(out, want['out']) < tol, rel_l2(out, want['out'])
        num = den = 0.0
        worst = ('', 0.0)
        for k, p in G.named_parameters():
            r = want['g_grads'][k]
            num += float((p.grad.double() - r).pow(2).sum()); den += float(r.pow(2).sum())
            if r.norm() > 1e-4:
                worst = max(worst, (k, rel_l2(p.grad, r)), key=lambda t: t[1])
        whole = (num / den) ** 0.5
        print('speech2gesture %s: out rel-L2 %.2e, whole gradient %.2e, worst tensor %s %.2e' % (mode, rel_l2(out, want['out']), whole, worst[0], worst[1]))
        # End-to-end gradients: this generator is 30 train-mode BatchNorm layers deep (statistics over 64 .. 8 704 rows at batch 32) and, at
        # these synthetic weights, amplifies the forward's round-off ~1e3-fold into the gradient: every block's backward agrees with fp64
        # autograd to 3e-7 when fed the same inputs (test_blockwise_backward_vs_fp64_autograd), the chained gradient to 5e-3 (fp32 mode,
        # measured; stock fp32 PyTorch on the same GPU: 2e-5) and only to ~0.16 in tf32 mode.  Held here: the fp32 figure; tf32 mode is held to
        # the mode's 1e-2 on poses and losses above and block by block in the test named.
        if mode == 'fp32':
            assert whole < 1e-2, whole
            assert worst[1] < 3e-2, worst
    finally:
        config.set_mode(old)


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-5), ('tf32', 5e-2)])      # tf32: the per-tensor bound of the fast mode (BatchNorm beta over 16 rows is the worst)
def test_blockwise_backward_vs_fp64_autograd(dev, mode, tol):
    """Every ConvNormRelu block of the generator (2-D and 1-D, SAME / VALID, stride 1 / 2) inside a REAL forward / backward sweep: the block's
    output, its input gradient, its convolution-weight gradient and its BatchNorm gamma / beta gradients against fp64 autograd of that block
    alone, fed the block's own recorded input and upstream gradient.  This is the well-conditioned form of gradient parity (see the note in
    test_train_iter_batch32_vs_fp64_oracle); the wiring between the blocks is checked against the reference golden by the emulated plan
    (tests/test_s2g_plan_emulated.py) and by the golden test above."""
    from tgb200 import config, s2g_engine
    old = config.set_mode(mode)
    rec = {}
    orig_fwd, orig_bwd = s2g_engine._ConvBlock.forward, s2g_engine._ConvBlock.backward

    def fwd(self, x, B_, H, W, training):
        rec[self.tag + '.in'] = x.clone()
        out = orig_fwd(self, x, B_, H, W, training)
        rec[self.tag + '.out'] = out[0].clone()
        return out

    def bwd(self, d, need_dx=True, param_grads=True):
        rec[self.tag + '.dout'] = d.clone()
        dx = orig_bwd(self, d, need_dx, param_grads)
        if dx is not None:
            rec[self.tag + '.dx'] = dx.clone()
        return dx
    try:
        s2g_engine._ConvBlock.forward, s2g_engine._ConvBlock.backward = fwd, bwd
        G, _, gsd, _ = _build(dev)
        G.train()
        Bb = 8
        spec, target = make_inputs(Bb, 41)
        ge = G.engine().ensure(dev)
        ge.forward(spec.to(dev), target[:, :N_PRE].to(dev), True)
        gen = torch.Generator().manual_seed(1)
        dout = (torch.randn(Bb, T, D, generator=gen) * 0.01).to(dev)
        ge.arena.zero_grad()
        ge.backward(dout.clone())
        torch.cuda.synchronize()
    finally:
        s2g_engine._ConvBlock.forward, s2g_engine._ConvBlock.backward = orig_fwd, orig_bwd
        config.set_mode(old)
    sd = {k: v.to(dev).double() if v.is_floating_point() else v.to(dev) for k, v in gsd.items()}
    pg = dict(G.named_parameters())
    worst = {}
    for blk in [ge.final] + ge.dec + ge.up + ge.down + ge.down1 + ge.first:
        Bq, H, W, cin, kh, kw, sh, sw, pt, pl, Ho, Wo, cout = blk.geom
        x = rec[blk.tag + '.in'].double().view(Bq, H, W, cin).permute(0, 3, 1, 2).clone().requires_grad_(True)
        w = sd[blk.conv + '.weight'].clone().requires_grad_(True)
        sdl = dict(sd)
        if blk.bn is not None:
            for q in ('.weight', '.bias'):
                sdl[blk.bn + q] = sd[blk.bn + q].clone().requires_grad_(True)
        m = ge.mod(blk.conv)
        pad = 'VALID' if (getattr(m, 'padding', 0) == 'VALID' or getattr(m, 'padding', 0) in (0, (0,), (0, 0))) else 'SAME'
        y = SO.conv_tf(x if w.dim() == 4 else x.squeeze(2), w, sd[blk.conv + '.bias'], m.stride, pad)
        if blk.bn is not None:
            y = F.leaky_relu(SO._bn(y, sdl, blk.bn, True, None), 0.2)
        yl = y if w.dim() == 4 else y.unsqueeze(2)
        errs = {'fwd': rel_l2(rec[blk.tag + '.out'].view(Bq, Ho, Wo, cout).permute(0, 3, 1, 2), yl)}
        (yl * rec[blk.tag + '.dout'].double().view(Bq, Ho, Wo, cout).permute(0, 3, 1, 2)).sum().backward()
        if blk.tag + '.dx' in rec:
            errs['dx'] = rel_l2(rec[blk.tag + '.dx'].view(Bq, H, W, cin).permute(0, 3, 1, 2), x.grad)
        errs['dW'] = rel_l2(pg[blk.conv + '.weight'].grad, w.grad)
        if blk.bn is not None:
            errs['dgamma'] = rel_l2(pg[blk.bn + '.weight'].grad, sdl[blk.bn + '.weight'].grad)
            errs['dbeta'] = rel_l2(pg[blk.bn + '.bias'].grad, sdl[blk.bn + '.bias'].grad)
        for k, v in errs.items():
            worst[k] = max(worst.get(k, ('', 0.0)), (blk.tag, v), key=lambda t: t[1])
            assert v < tol, (mode, blk.tag, k, v)
    print('speech2gesture blockwise (%s):' % mode, {k: (t, float('%.2e' % v)) for k, (t, v) in worst.items()})


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-5), ('tf32', 5e-2)])      # tf32: the per-tensor bound of the fast mode (BatchNorm beta over 16 rows is the worst)
def test_discriminator_fwd_bwd_vs_fp64_autograd(dev, mode, tol):
    """The discriminator alone (train-mode BatchNorm): scores, the gradient w.r.t. its input motion and every parameter gradient against
    fp64 autograd of the oracle on the SAME input (inside the training step its input is a second difference of generated poses, which
    amplifies the generator's round-off; here the comparison is well-conditioned)."""
    from tgb200 import config
    old = config.set_mode(mode)
    try:
        _, Dn, _, dsd = _build(dev)
        Dn.train()
        for Bb in (8, 32):
            gen = torch.Generator().manual_seed(5 + Bb)
            x = (torch.randn(Bb, T - 1, D, generator=gen) * 0.3).to(dev)
            de = Dn.engine().ensure(dev)
            de.arena.zero_grad()
            s = de.forward(x, True, slot='t')
            n = s.numel()
            gs = (torch.randn(n, generator=gen) * 0.1).to(dev)
            dp = de.backward(gs.clone().view(n, 1), need_dposes=True)
            from tgb200.engine import S_WGRAD, side
            side.join(S_WGRAD)
            torch.cuda.synchronize()
            f64 = {k: (v.to(dev).double() if v.is_floating_point() else v.to(dev)) for k, v in dsd.items()}
            params = {k: v.clone().requires_grad_(True) for k, v in f64.items() if v.is_floating_point() and 'running' not in k}
            full = dict(f64); full.update(params)
            xr = x.double().requires_grad_(True)
            out = SO.discriminator_forward(full, xr, True, {})
            assert rel_l2(s.view(Bb, -1), out.view(Bb, -1)) < tol
            (out.view(-1) * gs.double()).sum().backward()
            assert rel_l2(dp, xr.grad) < tol, rel_l2(dp, xr.grad)
            for k, p in Dn.named_parameters():
                if params[k].grad.norm() > 1e-6:
                    assert rel_l2(p.grad, params[k].grad) < tol, (mode, Bb, k, rel_l2(p.grad, params[k].grad))
    finally:
        config.set_mode(old)
