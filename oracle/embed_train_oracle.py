"""CPU oracle for training the FGD auto-encoder.  TEST INFRASTRUCTURE ONLY.

Functional restatement (CPU torch, fp32 / fp64) of
  * EmbeddingNet(mode='pose').forward in TRAIN mode - PoseEncoderConv + PoseDecoderConv with batch-statistics BatchNorm
    (scripts/model/embedding_net.py:42-82,165-217,276-308),
  * train_iter of scripts/train_feature_extractor.py:54-97 (L1 reconstruction + L1 of the frame differences, AE branch),
  * train_iter_embed / eval_embed of scripts/train_eval/train_joint_embed.py:5-65 restricted to the pose auto-encoder
    (no frame-difference term),
  * the Adam update of train_feature_extractor.py:134.
Pinned against the reference's own functions executed in the build container by oracle/make_golden_ae.py
(tests/golden/ae_train.npz, re-checked on every CPU run by tests/test_oracle_ae_golden.py).  Only tests/, smoke() and
bench.py's CPU legs may import this file; the product never does.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .trimodal_oracle import SD, Tensor, _is_param, _leafify, adam_step, batchnorm1d, leaky_relu

# parameters the reference leaves without a gradient in the AE branch (variational_encoding=False: logvar is computed, returned
# and never enters the loss, train_feature_extractor.py:60-86) - torch.optim.Adam skips them
UNUSED_PARAMS = ('pose_encoder.fc_logvar.weight', 'pose_encoder.fc_logvar.bias')
# parameters whose gradient is analytically ZERO: a constant per-channel shift that reaches a train-mode BatchNorm through linear
# maps only is removed by the mean subtraction (conv / linear biases in front of a BN, BN betas behind the identity
# 'LeakyReLU(True)', everything between out_net.6 and the decoder's first BN).  What autograd produces for them is fp32
# round-off (|g| ~ 1e-8), which Adam turns into +-lr steps: their post-step values - and, from the second step on, the running
# means that absorb them - are noise in the reference itself and are excluded from parity (cf. tests/test_oracle_golden.py).
ZERO_GRAD_PARAMS = ('pose_encoder.net.0.0.bias', 'pose_encoder.net.1.0.bias', 'pose_encoder.net.2.0.bias', 'pose_encoder.net.3.bias',
                    'pose_encoder.out_net.0.bias', 'pose_encoder.out_net.1.bias', 'pose_encoder.out_net.3.bias',
                    'pose_encoder.out_net.4.bias', 'pose_encoder.out_net.6.bias', 'pose_encoder.fc_mu.bias',
                    'decoder.pre_net.0.bias', 'decoder.net.0.bias', 'decoder.net.3.bias')
NOISY_RUNNING_MEANS = ('pose_encoder.net.0.1.running_mean', 'pose_encoder.net.1.1.running_mean', 'pose_encoder.net.2.1.running_mean',
                       'pose_encoder.out_net.1.running_mean', 'pose_encoder.out_net.4.running_mean', 'decoder.pre_net.1.running_mean',
                       'decoder.net.1.running_mean', 'decoder.net.4.running_mean')


def _lin(sd: SD, name: str, x: Tensor) -> Tensor:
    return x @ sd[name + '.weight'].t() + sd[name + '.bias']


def pose_encoder_conv(sd: SD, poses: Tensor, training: bool, stats: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """PoseEncoderConv.forward (embedding_net.py:67-82) -> (mu, logvar); z = mu when variational_encoding is False."""
    p = 'pose_encoder'
    x = poses.transpose(1, 2)                                                  # :69
    for i, stride in enumerate((1, 1, 2)):                                     # ConvNormRelu x3, :20-39,46-48
        x = F.conv1d(x, sd[f'{p}.net.{i}.0.weight'], sd[f'{p}.net.{i}.0.bias'], stride=stride)
        x = leaky_relu(batchnorm1d(x, sd, f'{p}.net.{i}.1', training, stats), 0.2)
    x = F.conv1d(x, sd[f'{p}.net.3.weight'], sd[f'{p}.net.3.bias'])          # :49
    x = x.flatten(1)                                                           # :71 (channel-major: index = c*12 + t)
    x = leaky_relu(batchnorm1d(_lin(sd, f'{p}.out_net.0', x), sd, f'{p}.out_net.1', training, stats), 1.0)   # LeakyReLU(True) = identity
    x = leaky_relu(batchnorm1d(_lin(sd, f'{p}.out_net.3', x), sd, f'{p}.out_net.4', training, stats), 1.0)
    x = _lin(sd, f'{p}.out_net.6', x)
    return _lin(sd, f'{p}.fc_mu', x), _lin(sd, f'{p}.fc_logvar', x)            # :75-76


def pose_decoder_conv(sd: SD, feat: Tensor, training: bool, stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """PoseDecoderConv.forward, length 34 (embedding_net.py:188-217)."""
    p = 'decoder'
    x = leaky_relu(batchnorm1d(_lin(sd, f'{p}.pre_net.0', feat), sd, f'{p}.pre_net.1', training, stats), 1.0)
    x = _lin(sd, f'{p}.pre_net.3', x)
    x = x.view(feat.shape[0], 4, -1)                                           # :213 (channel-major view [B,4,34])
    x = F.conv_transpose1d(x, sd[f'{p}.net.0.weight'], sd[f'{p}.net.0.bias'])
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.net.1', training, stats), 0.2)
    x = F.conv_transpose1d(x, sd[f'{p}.net.3.weight'], sd[f'{p}.net.3.bias'])
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.net.4', training, stats), 0.2)
    x = F.conv1d(x, sd[f'{p}.net.6.weight'], sd[f'{p}.net.6.bias'])
    x = F.conv1d(x, sd[f'{p}.net.7.weight'], sd[f'{p}.net.7.bias'])
    return x.transpose(1, 2)                                                   # :216


def embedding_net_pose(sd: SD, poses: Tensor, training: bool, stats: Optional[Dict[str, Tensor]] = None):
    """EmbeddingNet.forward(None, None, None, poses, None, variational_encoding=False) for mode='pose'
    -> (poses_feat, pose_mu, pose_logvar, out_poses) (embedding_net.py:276-308)."""
    mu, logvar = pose_encoder_conv(sd, poses, training, stats)
    return mu, mu, logvar, pose_decoder_conv(sd, mu, training, stats)


def recon_loss(recon: Tensor, target: Tensor, use_diff: bool) -> Tensor:
    """train_feature_extractor.py:64-72 (use_diff=True) / train_joint_embed.py:21-29 (use_diff=False): per-sample mean L1,
    plus the per-sample mean L1 of the frame-to-frame differences, summed over the batch."""
    loss = torch.mean(torch.abs(recon - target), dim=(1, 2))
    if use_diff:
        td = target[:, 1:] - target[:, :-1]
        rd = recon[:, 1:] - recon[:, :-1]
        loss = loss + torch.mean(torch.abs(rd - td), dim=(1, 2))
    return torch.sum(loss)


def eval_embed_oracle(sd: SD, target: Tensor) -> Tuple[float, Tensor]:
    """eval_embed(None, None, None, target, net) (train_joint_embed.py:54-65): batch mean of the per-sample mean L1; eval-mode BatchNorm."""
    _, _, _, recon = embedding_net_pose(sd, target, False)
    return torch.mean(torch.mean(torch.abs(recon - target), dim=(1, 2))).item(), recon


def train_iter_ae_oracle(sd: SD, opt: Dict[str, Dict[str, Tensor]], step_no: int, target: Tensor, lr: float, use_diff: bool,
                         betas=(0.5, 0.999), dtype=torch.float32):
    """One optimiser step of the auto-encoder: train_feature_extractor.train_iter (use_diff=True) or
    train_joint_embed.train_iter_embed on a mode='pose' net (use_diff=False).  opt = {'m': {...}, 'v': {...}} Adam moments
    (zeros before the first step), step_no = 1-based Adam step.  Returns loss, recon, every gradient, the updated
    state dict (parameters + BatchNorm buffers) and the new moments."""
    leaf = _leafify({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()})
    stats: Dict[str, Tensor] = {}
    feat, mu, logvar, recon = embedding_net_pose(leaf, target.to(dtype), True, stats)
    loss = recon_loss(recon, target.to(dtype), use_diff)
    keys = [k for k in leaf if _is_param(k)]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    g = {k: gr for k, gr in zip(keys, grads)}
    new_sd = {k: v.detach() for k, v in leaf.items()}
    new_sd.update(stats)
    new_m, new_v = dict(opt['m']), dict(opt['v'])
    for k in keys:
        if g[k] is None:                     # torch.optim.Adam skips parameters without a gradient
            assert k in UNUSED_PARAMS, k
            continue
        new_sd[k], new_m[k], new_v[k] = adam_step(leaf[k].detach(), g[k], opt['m'][k].to(dtype), opt['v'][k].to(dtype), step_no, lr,
                                                  betas[0], betas[1])
    grads_out = {k: (g[k] if g[k] is not None else torch.zeros_like(leaf[k])) for k in keys}
    return dict(loss=loss.item(), recon=recon.detach(), feat=feat.detach(), logvar=logvar.detach(), grads=grads_out, sd=new_sd,
                opt={'m': new_m, 'v': new_v})
