"""CPU oracle for the trimodal-gesture hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, written from scratch in functional torch (CPU,
fp32 or fp64), of the arithmetic that the reference's hot path executes
(SURVEY.md section 8a).  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``model/``, ``train_eval/``,
``tgb200/``) never does, and fails loudly when the CUDA library is missing.

Parity pinning: the reference ships no tests / golden vectors for this path
("parity unpinned" by the reference itself).  The oracle is
therefore pinned against the *reference modules themselves*, imported in the
build container by ``oracle/make_golden.py``; the resulting fixtures live in
``tests/golden`` and ``tests/test_oracle_golden.py`` re-checks the oracle
against them on every run (CPU, no reference needed).

Every function takes a plain ``dict[str, Tensor]`` that uses the reference's
``state_dict`` key names, so the same weights drive the reference, the oracle
and the CUDA implementation.  All stochastic inputs (dropout masks,
reparameterisation noise, the speaker permutation) are explicit arguments.

Reference anchors are ``scripts/...`` paths inside the reference repo.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class HotPathConfig:
    """Hyper-parameters of config/multimodal_context.yml:19-40 (defaults = that file)."""
    n_poses: int = 34
    n_pre_poses: int = 4
    pose_dim: int = 27
    hidden_size: int = 300
    n_layers: int = 4
    dropout_prob: float = 0.3
    emb_dropout: float = 0.1
    wordembed_dim: int = 300
    n_words: int = 20000
    n_speakers: int = 1371          # z_obj.n_words
    z_size: int = 16
    audio_len: int = 36267          # lmdb_data_loader.py:68
    d_hidden: int = 64              # multimodal_context_net.py:212
    d_layers: int = 4
    # train_gan.py hyper-parameters
    loss_warmup: int = 10
    loss_gan_weight: float = 5.0
    loss_regression_weight: float = 500.0
    loss_kld_weight: float = 0.1
    loss_reg_weight: float = 0.05
    learning_rate: float = 5e-4
    discriminator_lr_weight: float = 0.2

    @property
    def gru_in(self) -> int:        # multimodal_context_net.py:73,85-86
        return 32 + 32 + self.pose_dim + 1 + self.z_size


# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
def leaky_relu(x: Tensor, slope: float) -> Tensor:
    """nn.LeakyReLU(slope).  NB nn.LeakyReLU(True) == slope 1.0 == identity
    (multimodal_context_net.py:102,216,219; embedding_net.py:57,60)."""
    return torch.where(x >= 0, x, x * slope)


def batchnorm1d(x: Tensor, sd: SD, prefix: str, training: bool,
                stats_out: Optional[Dict[str, Tensor]] = None, momentum: float = 0.1,
                eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm1d on [B,C,L] or [B,C].  Train: biased batch variance for the
    normalisation, running stats updated with the UNBIASED variance (torch semantics,
    probed in SURVEY.md 8c).  ``stats_out`` receives the new running buffers."""
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    dims = (0, 2) if x.dim() == 3 else (0,)
    shape = (1, -1, 1) if x.dim() == 3 else (1, -1)
    if training:
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if stats_out is not None:
            n = x.numel() // x.shape[1]
            unbiased = var * (n / max(n - 1, 1))
            rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
            rm = stats_out.get(prefix + '.running_mean', rm)
            rv = stats_out.get(prefix + '.running_var', rv)
            nbt = stats_out.get(prefix + '.num_batches_tracked', sd[prefix + '.num_batches_tracked'])
            stats_out[prefix + '.running_mean'] = ((1 - momentum) * rm + momentum * mean).detach()
            stats_out[prefix + '.running_var'] = ((1 - momentum) * rv + momentum * unbiased).detach()
            stats_out[prefix + '.num_batches_tracked'] = nbt + 1
    else:
        mean, var = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    xhat = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps)
    return xhat * w.view(shape) + b.view(shape)


def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, norm over dims (1,2) per
    output channel (tcn.py:19-25)."""
    norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
    return v * (g / norm)


def gru_cell_sequence(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                      reverse: bool) -> Tensor:
    """One direction of one nn.GRU layer, batch_first, h0 = 0.  Gate row order r,z,n
    (torch semantics probed in SURVEY.md 8a row 5):
        r = s(Wir x + bir + Whr h + bhr); z = s(Wiz x + biz + Whz h + bhz)
        n = tanh(Win x + bin + r*(Whn h + bhn)); h' = (1-z)*n + z*h
    """
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih                      # [B,T,3H] batched input projection
    h = x.new_zeros(B, H)
    outs: List[Optional[Tensor]] = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gh = h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi[:, t].split(H, dim=1)
        h_r, h_z, h_n = gh.split(H, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)               # [B,T,H]


def gru_bidirectional(x: Tensor, sd: SD, prefix: str, n_layers: int,
                      layer_masks: Optional[List[Optional[Tensor]]] = None) -> Tensor:
    """nn.GRU(num_layers, bidirectional=True, batch_first=True).  ``layer_masks[l]`` is the
    multiplicative inter-layer dropout mask (already scaled by 1/(1-p)) applied to the
    OUTPUT of layer l for l < n_layers-1 (torch applies no dropout after the last layer);
    None == eval mode / p = 0.  Returns [B,T,2H] (fwd || rev)."""
    inp = x
    for l in range(n_layers):
        outs = []
        for suffix, rev in (('', False), ('_reverse', True)):
            outs.append(gru_cell_sequence(
                inp, sd[f'{prefix}.weight_ih_l{l}{suffix}'], sd[f'{prefix}.weight_hh_l{l}{suffix}'],
                sd[f'{prefix}.bias_ih_l{l}{suffix}'], sd[f'{prefix}.bias_hh_l{l}{suffix}'], rev))
        inp = torch.cat(outs, dim=2)
        if layer_masks is not None and l < n_layers - 1 and layer_masks[l] is not None:
            inp = inp * layer_masks[l]
    return inp


# --------------------------------------------------------------------------------------
# PoseGenerator pieces
# --------------------------------------------------------------------------------------
def wav_encoder(sd: SD, prefix: str, wav: Tensor, training: bool,
                stats_out: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """WavEncoder.forward (multimodal_context_net.py:9-28): [B,L] -> [B,34,32]."""
    p = prefix + '.feat_extractor'
    x = wav.unsqueeze(1)
    x = F.conv1d(x, sd[p + '.0.weight'], sd[p + '.0.bias'], stride=5, padding=1600)
    x = leaky_relu(batchnorm1d(x, sd, p + '.1', training, stats_out), 0.3)
    x = F.conv1d(x, sd[p + '.3.weight'], sd[p + '.3.bias'], stride=6)
    x = leaky_relu(batchnorm1d(x, sd, p + '.4', training, stats_out), 0.3)
    x = F.conv1d(x, sd[p + '.6.weight'], sd[p + '.6.bias'], stride=6)
    x = leaky_relu(batchnorm1d(x, sd, p + '.7', training, stats_out), 0.3)
    x = F.conv1d(x, sd[p + '.9.weight'], sd[p + '.9.bias'], stride=6)
    return x.transpose(1, 2)


def causal_conv(x: Tensor, w: Tensor, b: Tensor, dilation: int) -> Tensor:
    """Conv1d(pad=(k-1)*d both sides) followed by Chomp1d(pad) == left-pad only
    (tcn.py:7-13,19-31).  x: [B,C,T]."""
    k = w.shape[2]
    return F.conv1d(F.pad(x, ((k - 1) * dilation, 0)), w, b, dilation=dilation)


def text_encoder_tcn(sd: SD, prefix: str, in_text: Tensor, n_layers: int,
                     masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """TextEncoderTCN.forward (multimodal_context_net.py:57-61) + TemporalConvNet
    (tcn.py:43-64).  masks: {'emb': [B,T,E], 'tcn{i}_{1|2}': [B,C,T]} multiplicative dropout
    masks (scaled); None == eval."""
    emb = sd[prefix + '.embedding.weight'][in_text]            # [B,T,E]
    if masks is not None and 'emb' in masks:
        emb = emb * masks['emb']
    x = emb.transpose(1, 2)
    for i in range(n_layers):
        q = f'{prefix}.tcn.network.{i}'
        d = 2 ** i
        w1 = weight_norm_weight(sd[q + '.conv1.weight_g'], sd[q + '.conv1.weight_v'])
        w2 = weight_norm_weight(sd[q + '.conv2.weight_g'], sd[q + '.conv2.weight_v'])
        y = torch.relu(causal_conv(x, w1, sd[q + '.conv1.bias'], d))
        if masks is not None and f'tcn{i}_1' in masks:
            y = y * masks[f'tcn{i}_1']
        y = torch.relu(causal_conv(y, w2, sd[q + '.conv2.bias'], d))
        if masks is not None and f'tcn{i}_2' in masks:
            y = y * masks[f'tcn{i}_2']
        if (q + '.downsample.weight') in sd:                    # only when channels differ (tcn.py:33)
            res = F.conv1d(x, sd[q + '.downsample.weight'], sd[q + '.downsample.bias'])
        else:
            res = x
        x = torch.relu(y + res)
    y = x.transpose(1, 2) @ sd[prefix + '.decoder.weight'].t() + sd[prefix + '.decoder.bias']
    return y


def speaker_style(sd: SD, vid: Tensor, eps: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """multimodal_context_net.py:125-131 + embedding_net.reparameterize :10-13."""
    e = sd['speaker_embedding.0.weight'][vid]
    e = e @ sd['speaker_embedding.1.weight'].t() + sd['speaker_embedding.1.bias']
    mu = e @ sd['speaker_mu.weight'].t() + sd['speaker_mu.bias']
    logvar = e @ sd['speaker_logvar.weight'].t() + sd['speaker_logvar.bias']
    z = mu + eps * torch.exp(0.5 * logvar)
    return z, mu, logvar


def pose_generator_forward(sd: SD, cfg: HotPathConfig, pre_seq: Tensor, in_text: Tensor,
                           in_audio: Tensor, vid: Tensor, eps: Tensor, training: bool = False,
                           masks: Optional[Dict[str, Tensor]] = None,
                           stats_out: Optional[Dict[str, Tensor]] = None,
                           input_context: str = 'both', z_mode: Optional[str] = 'speaker'):
    """PoseGenerator.forward (multimodal_context_net.py:110-160).  Returns (poses [B,T,D], z, mu, logvar).
    input_context in {'both','audio','text','none'} (:139-148); z_mode 'speaker' (z_obj is a Vocab: embedding -> mu/logvar ->
    reparameterize with eps, :125-131), 'random' (any other truthy z_obj: z = eps ~ N(0,1), :132-134) or None (:135-137).
    masks keys: 'emb', 'tcn{i}_{1|2}', 'gru{l}' (l < n_layers-1)."""
    parts = [pre_seq]
    if input_context in ('both', 'audio'):
        parts.append(wav_encoder(sd, 'audio_encoder', in_audio, training, stats_out))
    if input_context in ('both', 'text'):
        parts.append(text_encoder_tcn(sd, 'text_encoder', in_text, cfg.n_layers, masks))
    z = mu = logvar = None
    if z_mode == 'speaker':
        z, mu, logvar = speaker_style(sd, vid, eps)
    elif z_mode == 'random':
        z = eps
    if z is not None:
        parts.append(z.unsqueeze(1).expand(-1, pre_seq.shape[1], -1))
    in_data = torch.cat(parts, dim=2)
    gmasks = None
    if masks is not None:
        gmasks = [masks.get(f'gru{l}') for l in range(cfg.n_layers)]
    out = gru_bidirectional(in_data, sd, 'gru', cfg.n_layers, gmasks)
    H = cfg.hidden_size
    out = out[:, :, :H] + out[:, :, H:]
    y = out.reshape(-1, H) @ sd['out.0.weight'].t() + sd['out.0.bias']
    y = leaky_relu(y, 1.0)                                       # nn.LeakyReLU(True) == identity
    y = y @ sd['out.2.weight'].t() + sd['out.2.bias']
    return y.reshape(pre_seq.shape[0], pre_seq.shape[1], -1), z, mu, logvar


# --------------------------------------------------------------------------------------
# ConvDiscriminator
# --------------------------------------------------------------------------------------
def conv_discriminator_forward(sd: SD, cfg: HotPathConfig, poses: Tensor, training: bool = False,
                               masks: Optional[Dict[str, Tensor]] = None,
                               stats_out: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ConvDiscriminator.forward (multimodal_context_net.py:232-252): [B,34,27] -> [B,1]."""
    x = poses.transpose(1, 2)
    x = F.conv1d(x, sd['pre_conv.0.weight'], sd['pre_conv.0.bias'])
    x = leaky_relu(batchnorm1d(x, sd, 'pre_conv.1', training, stats_out), 1.0)
    x = F.conv1d(x, sd['pre_conv.3.weight'], sd['pre_conv.3.bias'])
    x = leaky_relu(batchnorm1d(x, sd, 'pre_conv.4', training, stats_out), 1.0)
    x = F.conv1d(x, sd['pre_conv.6.weight'], sd['pre_conv.6.bias'])
    x = x.transpose(1, 2)
    gmasks = None
    if masks is not None:
        gmasks = [masks.get(f'gru{l}') for l in range(cfg.d_layers)]
    out = gru_bidirectional(x, sd, 'gru', cfg.d_layers, gmasks)
    H = cfg.d_hidden
    out = out[:, :, :H] + out[:, :, H:]
    B = poses.shape[0]
    y = out.reshape(-1, H) @ sd['out.weight'].t() + sd['out.bias']
    y = y.view(B, -1) @ sd['out2.weight'].t() + sd['out2.bias']
    return torch.sigmoid(y)


# --------------------------------------------------------------------------------------
# losses (train_gan.py:41,53-56,67-82)
# --------------------------------------------------------------------------------------
def huber(x: Tensor, y: Tensor, beta: float) -> Tensor:
    """F.smooth_l1_loss(x/beta, y/beta, reduction='none') * beta with torch's own beta=1."""
    d = (x / beta - y / beta).abs()
    return torch.where(d < 1, 0.5 * d * d, d - 0.5) * beta


def make_pre_seq(target: Tensor, n_pre: int) -> Tensor:
    """train_gan.py:20-22."""
    pre = target.new_zeros(target.shape[0], target.shape[1], target.shape[2] + 1)
    pre[:, :n_pre, :-1] = target[:, :n_pre]
    pre[:, :n_pre, -1] = 1
    return pre


def dis_loss(d_real: Tensor, d_fake: Tensor) -> Tensor:
    return torch.sum(-torch.mean(torch.log(d_real + 1e-8) + torch.log(1 - d_fake + 1e-8)))


def gen_losses(cfg: HotPathConfig, out: Tensor, target: Tensor, d_out: Tensor, out_rand: Tensor,
               z: Tensor, z_rand: Tensor, mu: Tensor, logvar: Tensor, after_warmup: bool):
    hub = huber(out, target, 0.1).mean()
    gen = -torch.mean(torch.log(d_out + 1e-8))
    pose_l1 = huber(out, out_rand.detach(), 0.05).sum(dim=1).sum(dim=1)
    z_l1 = (z.detach() - z_rand.detach()).abs().mean(dim=1)
    div = torch.clamp(-(pose_l1 / (z_l1 + 1.0e-5)), min=-1000).mean()
    kld = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
    loss = cfg.loss_regression_weight * hub + cfg.loss_kld_weight * kld + cfg.loss_reg_weight * div
    if after_warmup:
        loss = loss + cfg.loss_gan_weight * gen
    return loss, hub, gen, div, kld


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam, betas (0.5,0.999), eps 1e-8, no weight decay; train.py:104-109)
# --------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.5, b2: float = 0.999, eps: float = 1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


# --------------------------------------------------------------------------------------
# one G+D iteration (train_gan.py:13-103), fully functional
# --------------------------------------------------------------------------------------
@dataclass
class StepNoise:
    """Every random draw of one train_iter_gan call, in call order: three G forwards
    (D-step, G-step, G-step with permuted speakers) and three D forwards (real, fake, gen)."""
    eps: List[Tensor]                                   # 3 x [B,16]
    perm: Tensor                                        # [B] int64 (torch.randperm, train_gan.py:62)
    g_masks: List[Optional[Dict[str, Tensor]]] = field(default_factory=lambda: [None, None, None])
    d_masks: List[Optional[Dict[str, Tensor]]] = field(default_factory=lambda: [None, None, None])


def _leafify(sd: SD, buffers_suffix=('running_mean', 'running_var', 'num_batches_tracked')) -> SD:
    out = {}
    for k, v in sd.items():
        if k.endswith(buffers_suffix) or not v.is_floating_point():
            out[k] = v.clone()
        else:
            out[k] = v.detach().clone().requires_grad_(True)
    return out


def _is_param(k: str) -> bool:
    return not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))


def train_iter_gan_oracle(cfg: HotPathConfig, epoch: int, g_sd: SD, d_sd: SD,
                          g_opt: Dict[str, Dict[str, Tensor]], d_opt: Dict[str, Dict[str, Tensor]],
                          step_no: int, in_text: Tensor, in_audio: Tensor, target: Tensor,
                          vid: Tensor, noise: StepNoise):
    """Functional restatement of train_iter_gan.  g_opt/d_opt: {'m': {k: T}, 'v': {k: T}}.
    Returns dict(losses=..., g_sd=..., d_sd=..., g_grads=..., d_grads=..., g_opt=..., d_opt=...).
    d_grads are those used for D's update (None during warm-up)."""
    after = epoch > cfg.loss_warmup
    do_d = after and cfg.loss_gan_weight > 0.0
    pre_seq = make_pre_seq(target, cfg.n_pre_poses)
    g_stats: Dict[str, Tensor] = {}
    d_stats: Dict[str, Tensor] = {}
    g = _leafify(g_sd)
    d = _leafify(d_sd)
    ret: Dict[str, object] = {}
    losses: Dict[str, float] = {}
    d_grads = None
    dis_error = None

    def with_stats(sd, stats):
        merged = dict(sd)
        merged.update(stats)
        return merged

    if do_d:
        out0, *_ = pose_generator_forward(with_stats(g, g_stats), cfg, pre_seq, in_text, in_audio, vid,
                                          noise.eps[0], True, noise.g_masks[0], g_stats)
        d_real = conv_discriminator_forward(with_stats(d, d_stats), cfg, target, True, noise.d_masks[0], d_stats)
        d_fake = conv_discriminator_forward(with_stats(d, d_stats), cfg, out0.detach(), True, noise.d_masks[1], d_stats)
        dis_error = dis_loss(d_real, d_fake)
        keys = [k for k in d if _is_param(k)]
        grads = torch.autograd.grad(dis_error, [d[k] for k in keys], allow_unused=True)
        d_grads = {k: (gr if gr is not None else torch.zeros_like(d[k])) for k, gr in zip(keys, grads)}
        lr_d = cfg.learning_rate * cfg.discriminator_lr_weight
        new_d, new_m, new_v = {}, {}, {}
        for k in keys:
            p2, m2, v2 = adam_step(d[k].detach(), d_grads[k], d_opt['m'][k], d_opt['v'][k], step_no, lr_d)
            new_d[k], new_m[k], new_v[k] = p2, m2, v2
        d_opt = {'m': new_m, 'v': new_v}
        d = _leafify(with_stats({**d, **new_d}, d_stats))
    # ---- G step
    out, z, mu, logvar = pose_generator_forward(with_stats(g, g_stats), cfg, pre_seq, in_text, in_audio, vid,
                                                noise.eps[1], True, noise.g_masks[1], g_stats)
    d_out = conv_discriminator_forward(with_stats(d, d_stats), cfg, out, True, noise.d_masks[2], d_stats)
    rand_vid = vid[noise.perm]
    out_r, z_r, _, _ = pose_generator_forward(with_stats(g, g_stats), cfg, pre_seq, in_text, in_audio, rand_vid,
                                              noise.eps[2], True, noise.g_masks[2], g_stats)
    loss, hub, gen, div, kld = gen_losses(cfg, out, target, d_out, out_r, z, z_r, mu, logvar, after)
    gkeys = [k for k in g if _is_param(k)]
    grads = torch.autograd.grad(loss, [g[k] for k in gkeys], allow_unused=True)
    g_grads = {k: (gr if gr is not None else torch.zeros_like(g[k])) for k, gr in zip(gkeys, grads)}
    new_g, new_m, new_v = {}, {}, {}
    for k in gkeys:
        p2, m2, v2 = adam_step(g[k].detach(), g_grads[k], g_opt['m'][k], g_opt['v'][k], step_no, cfg.learning_rate)
        new_g[k], new_m[k], new_v[k] = p2, m2, v2
    g_opt = {'m': new_m, 'v': new_v}
    losses['loss'] = cfg.loss_regression_weight * hub.item()
    losses['KLD'] = cfg.loss_kld_weight * kld.item()
    losses['DIV_REG'] = cfg.loss_reg_weight * div.item()
    if do_d:
        losses['gen'] = cfg.loss_gan_weight * gen.item()
        losses['dis'] = dis_error.item()
    g_final = {k: v.detach() for k, v in with_stats({**g, **new_g}, g_stats).items()}
    d_final = {k: v.detach() for k, v in with_stats(d, d_stats).items()}
    ret.update(losses=losses, g_sd=g_final, d_sd=d_final, g_grads=g_grads, d_grads=d_grads,
               g_opt=g_opt, d_opt=d_opt, out=out.detach(), total_loss=loss.item())
    return ret


# --------------------------------------------------------------------------------------
# EmbeddingNet (mode='pose') + FGD (embedding_net.py:42-82,165-217; embedding_space_evaluator.py)
# --------------------------------------------------------------------------------------
def pose_encoder_conv(sd: SD, poses: Tensor) -> Tensor:
    """PoseEncoderConv.forward, eval mode, variational_encoding=False -> feature = mu."""
    p = 'pose_encoder'
    x = poses.transpose(1, 2)
    for i, (stride,) in enumerate(((1,), (1,), (2,))):
        x = F.conv1d(x, sd[f'{p}.net.{i}.0.weight'], sd[f'{p}.net.{i}.0.bias'], stride=stride)
        x = leaky_relu(batchnorm1d(x, sd, f'{p}.net.{i}.1', False), 0.2)
    x = F.conv1d(x, sd[f'{p}.net.3.weight'], sd[f'{p}.net.3.bias'])
    x = x.flatten(1)
    x = x @ sd[f'{p}.out_net.0.weight'].t() + sd[f'{p}.out_net.0.bias']
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.out_net.1', False), 1.0)
    x = x @ sd[f'{p}.out_net.3.weight'].t() + sd[f'{p}.out_net.3.bias']
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.out_net.4', False), 1.0)
    x = x @ sd[f'{p}.out_net.6.weight'].t() + sd[f'{p}.out_net.6.bias']
    mu = x @ sd[f'{p}.fc_mu.weight'].t() + sd[f'{p}.fc_mu.bias']
    return mu


def pose_decoder_conv(sd: SD, feat: Tensor) -> Tensor:
    """PoseDecoderConv.forward (length 34), eval mode."""
    p = 'decoder'
    x = feat @ sd[f'{p}.pre_net.0.weight'].t() + sd[f'{p}.pre_net.0.bias']
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.pre_net.1', False), 1.0)
    x = x @ sd[f'{p}.pre_net.3.weight'].t() + sd[f'{p}.pre_net.3.bias']
    x = x.view(feat.shape[0], 4, -1)
    x = F.conv_transpose1d(x, sd[f'{p}.net.0.weight'], sd[f'{p}.net.0.bias'])
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.net.1', False), 0.2)
    x = F.conv_transpose1d(x, sd[f'{p}.net.3.weight'], sd[f'{p}.net.3.bias'])
    x = leaky_relu(batchnorm1d(x, sd, f'{p}.net.4', False), 0.2)
    x = F.conv1d(x, sd[f'{p}.net.6.weight'], sd[f'{p}.net.6.bias'])
    x = F.conv1d(x, sd[f'{p}.net.7.weight'], sd[f'{p}.net.7.bias'])
    return x.transpose(1, 2)


def embedding_net_pose_forward(sd: SD, poses: Tensor) -> Tuple[Tensor, Tensor]:
    """EmbeddingNet.forward(None, None, pre, poses, 'pose') -> (poses_feat, out_poses)."""
    feat = pose_encoder_conv(sd, poses)
    return feat, pose_decoder_conv(sd, feat)


def sqrtm_psd_product(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy.linalg.sqrtm(a @ b) (embedding_space_evaluator.py:138)."""
    from scipy import linalg
    res = linalg.sqrtm(a.dot(b))
    if isinstance(res, tuple):
        res = res[0]
    return res


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    """calculate_frechet_distance (embedding_space_evaluator.py:103-156)."""
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    covmean = sqrtm_psd_product(sigma1, sigma2)
    if not np.isfinite(covmean).all():
        off = np.eye(sigma1.shape[0]) * eps
        covmean = sqrtm_psd_product(sigma1 + off, sigma2 + off)
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError('Imaginary component {}'.format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def fgd_scores(generated_feats: np.ndarray, real_feats: np.ndarray) -> Tuple[float, float]:
    """EmbeddingSpaceEvaluator.get_scores (embedding_space_evaluator.py:74-101)."""
    a_mu, a_sigma = np.mean(generated_feats, axis=0), np.cov(generated_feats, rowvar=False)
    b_mu, b_sigma = np.mean(real_feats, axis=0), np.cov(real_feats, rowvar=False)
    try:
        fd = frechet_distance(a_mu, a_sigma, b_mu, b_sigma)
    except ValueError:
        fd = 1e+10
    feat_dist = float(np.mean(np.sum(np.abs(real_feats - generated_feats), axis=1)))
    return fd, feat_dist
