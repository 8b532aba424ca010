"""train_iter_speech2gesture (reference: scripts/train_eval/train_speech2gesture.py:5-37; call site train.py:199-201) as a launch plan over the
C ABI: generator forward, the three discriminator passes (real / fake for D's LSGAN step, the generated motion again for G's step - after D's
Adam update, with train-mode BatchNorm statistics each time), L1 + LSGAN losses with their gradients in the same kernels, both backward sweeps
and both flat Adam updates.  Same signature, same returned dict.  CUDA tensors only."""
from typing import Dict

import torch

from tgb200 import _lib, ops
from tgb200.engine import S_WGRAD, side


def _unwrap(m):
    return m.module if hasattr(m, 'module') and not hasattr(m, 'engine') else m


def train_iter_speech2gesture(args, in_spec, target_poses, pose_decoder, discriminator, pose_dec_optim, dis_optim, loss_fn=None) -> Dict[str, float]:
    _lib.require_cuda()
    G, D = _unwrap(pose_decoder), _unwrap(discriminator)
    if not target_poses.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('train_iter_speech2gesture runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
    assert loss_fn is None or isinstance(loss_fn, torch.nn.L1Loss), 'train.py:62 constructs torch.nn.L1Loss() for this model'
    dev = target_poses.device
    ge, de = G.engine().ensure(dev), D.engine().ensure(dev)
    ws = ge.ws
    target = target_poses.contiguous().float()
    B, T, Dm = target.shape
    w_reg, w_gan = float(args.loss_regression_weight), float(args.loss_gan_weight)

    # ---- generation (train_speech2gesture.py:8-13)
    pre = target[:, 0:args.n_pre_poses]
    out = ge.forward(in_spec, pre, G.training)
    tmotion = ws.get('s2g.tmotion', (B, T - 1, Dm)); omotion = ws.get('s2g.omotion', (B, T - 1, Dm))
    ops.time_diff_fwd(target, tmotion, B, T, Dm)
    ops.time_diff_fwd(out, omotion, B, T, Dm)
    sc = ws.get('s2g.scalars', (4,), torch.float64); sc.zero_()

    # ---- train D (:17-24): mse(1, D(real)) + mse(0, D(fake.detach()))
    de.arena.zero_grad()
    s_real = de.forward(tmotion, D.training, slot='r'); ctx_r = de.ctx
    n = s_real.numel()
    g_real = ws.get('s2g.g_real', (n,))
    ops.mse_const(s_real, n, 1.0, 1.0, sc[0:], g_real)
    s_fake = de.forward(omotion, D.training, slot='f'); ctx_f = de.ctx
    g_fake = ws.get('s2g.g_fake', (n,))
    ops.mse_const(s_fake, n, 0.0, 1.0, sc[0:], g_fake)
    de.backward(g_real.view(n, 1), need_dposes=False, ctx=ctx_r)
    de.backward(g_fake.view(n, 1), need_dposes=False, ctx=ctx_f)
    side.join(S_WGRAD)          # bias column sums of the tensor-core weight-gradient path run on their own stream
    de.arena.adam_step(dis_optim)

    # ---- train G (:28-35): w_reg * L1(out, target) + w_gan * mse(1, D(fake))
    ge.arena.zero_grad()
    d_out = ws.get('s2g.d_out', (B, T, Dm))
    ops.l1_loss(out, target, B * T * Dm, w_reg, sc[1:], d_out)
    s_gen = de.forward(omotion, D.training, slot='g')
    g_gen = ws.get('s2g.g_gen', (n,))
    ops.mse_const(s_gen, n, 1.0, w_gan, sc[2:], g_gen)
    # (the reference also accumulates this loss's gradient into D's parameters; dis_optim.zero_grad() discards it at the next call)
    d_om = de.backward(g_gen.view(n, 1), need_dposes=True, param_grads=False)
    ops.time_diff_bwd(d_om, d_out, B, T, Dm, accumulate=True)
    ge.backward(d_out)
    side.join(S_WGRAD)
    ge.arena.adam_step(pose_dec_optim)

    s = sc.cpu().tolist()
    return {'loss': w_reg * s[1], 'gen': w_gan * s[2], 'dis': s[0]}
