"""CPU oracle for the joint-embedding model (second half).  TEST INFRASTRUCTURE ONLY.

Functional restatement (CPU torch, fp32 / fp64) of EmbeddingNet(mode != 'pose') - ContextEncoder (TextEncoderTCN + WavEncoder +
2-layer unidirectional GRU H=256 + MLP + reparameterised 32-d latent, scripts/model/embedding_net.py:220-259), PoseEncoderConv (:42-82)
and PoseDecoderGRU (:130-162) - and of train_iter_embed / eval_embed (scripts/train_eval/train_joint_embed.py:5-65) for the two
branches `input_mode` can resolve to ('speech': decode the context latent, 'pose': decode the pose latent; 'random' flips a Python
coin between them, embedding_net.py:295-296).  Pinned against the reference modules executed in the build container by
oracle/make_golden_joint.py (tests/golden/joint_embed.npz).  Only tests/ may import this file."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .embed_train_oracle import pose_encoder_conv
from .trimodal_oracle import (SD, Tensor, _is_param, _leafify, adam_step, batchnorm1d, gru_bidirectional, gru_cell_sequence,
                              text_encoder_tcn, wav_encoder)


def _lin(sd: SD, name: str, x: Tensor) -> Tensor:
    return x @ sd[name + '.weight'].t() + sd[name + '.bias']


def context_encoder(sd: SD, in_text: Tensor, in_audio: Tensor, eps: Tensor, training: bool, n_tcn_layers: int,
                    masks: Optional[Dict[str, Tensor]] = None, stats: Optional[Dict[str, Tensor]] = None):
    """ContextEncoder.forward (embedding_net.py:244-259) -> (z, mu, logvar).  z = mu + eps*exp(logvar/2) is drawn in EVERY mode
    (the reparameterisation is unconditional, :258); eps is explicit here."""
    p = 'context_encoder'
    text = text_encoder_tcn(sd, p + '.text_encoder', in_text, n_tcn_layers, masks)            # [B,34,32]
    audio = wav_encoder(sd, p + '.audio_encoder', in_audio, training, stats)                 # [B,34,32]
    x = torch.cat((audio, text), dim=2)                                                      # :250
    for l in range(2):                                                                       # nn.GRU(64, 256, num_layers=2), :227-228
        x = gru_cell_sequence(x, sd[f'{p}.gru.weight_ih_l{l}'], sd[f'{p}.gru.weight_hh_l{l}'], sd[f'{p}.gru.bias_ih_l{l}'],
                              sd[f'{p}.gru.bias_hh_l{l}'], False)
    h = x[:, -1]                                                                             # :253
    h = torch.relu(batchnorm1d(_lin(sd, p + '.out.0', h), sd, p + '.out.1', training, stats))
    h = _lin(sd, p + '.out.3', h)
    mu, logvar = _lin(sd, p + '.fc_mu', h), _lin(sd, p + '.fc_logvar', h)
    return mu + eps * torch.exp(0.5 * logvar), mu, logvar


def pose_decoder_gru(sd: SD, latent: Tensor, pre_poses: Tensor, training: bool, gen_length: int,
                     gru_masks: Optional[List[Optional[Tensor]]] = None, stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """PoseDecoderGRU.forward (embedding_net.py:151-162): latent [B,32], pre_poses [B,4,D] -> [B,gen_length,D]."""
    p = 'decoder'
    B = pre_poses.shape[0]
    f = _lin(sd, p + '.pre_pose_net.0', pre_poses.reshape(B, -1))
    f = torch.relu(batchnorm1d(f, sd, p + '.pre_pose_net.1', training, stats))
    f = _lin(sd, p + '.pre_pose_net.3', f)
    feat = torch.cat((f, latent), dim=1).unsqueeze(1).repeat(1, gen_length, 1)               # :153-154
    out = gru_bidirectional(feat, sd, p + '.gru', 4, gru_masks)
    H = out.shape[2] // 2
    out = out[:, :, :H] + out[:, :, H:]                                                      # :157
    out = _lin(sd, p + '.out.2', _lin(sd, p + '.out.0', out))                                # LeakyReLU(True) == identity, :147
    return out


def embedding_net_joint(sd: SD, in_text, in_audio, pre_poses, poses, input_mode: str, eps: Optional[Tensor], training: bool,
                        n_tcn_layers: int = 4, masks=None, gru_masks=None, stats=None):
    """EmbeddingNet.forward (embedding_net.py:276-308), variational_encoding=False, input_mode in {'speech','pose'}
    -> (context_feat, context_mu, context_logvar, poses_feat, pose_mu, pose_logvar, out_poses)."""
    c_feat = c_mu = c_lv = None
    if in_text is not None and in_audio is not None:
        c_feat, c_mu, c_lv = context_encoder(sd, in_text, in_audio, eps, training, n_tcn_layers, masks, stats)
    p_feat = p_mu = p_lv = None
    if poses is not None:
        p_mu, p_lv = pose_encoder_conv(sd, poses, training, stats)
        p_feat = p_mu
    latent = c_feat if input_mode == 'speech' else p_feat
    out = pose_decoder_gru(sd, latent, pre_poses, training, poses.shape[1] if poses is not None else 34, gru_masks, stats)
    return c_feat, c_mu, c_lv, p_feat, p_mu, p_lv, out


def train_iter_embed_oracle(sd: SD, opt, step_no: Dict[str, int], in_text, in_audio, target, n_pre: int, input_mode: str, eps, lr: float,
                            masks=None, gru_masks=None, betas=(0.5, 0.999), dtype=torch.float32, n_tcn_layers: int = 4):
    """train_iter_embed (train_joint_embed.py:5-51), variational_encoding=False: loss = sum_b mean|recon - target|; only the
    parameters on the chosen branch receive a gradient, torch.optim.Adam skips the others (their moments AND step counts stay put):
    step_no maps a parameter name to the number of Adam updates it has seen so far."""
    leaf = _leafify({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()})
    stats: Dict[str, Tensor] = {}
    cast = lambda t: None if t is None else (t.to(dtype) if t.is_floating_point() else t)
    tgt = cast(target)
    pre = tgt[:, :n_pre]                                                                     # :6
    outs = embedding_net_joint(leaf, in_text, cast(in_audio), pre, tgt, input_mode, cast(eps), True, n_tcn_layers,
                               None if masks is None else {k: cast(v) for k, v in masks.items()},
                               None if gru_masks is None else [cast(m) for m in gru_masks], stats)
    recon = outs[6]
    loss = torch.sum(torch.mean(torch.abs(recon - tgt), dim=(1, 2)))                         # :21-29,46
    keys = [k for k in leaf if _is_param(k)]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    new_sd = {k: v.detach() for k, v in leaf.items()}
    new_sd.update(stats)
    new_m, new_v, new_step = dict(opt['m']), dict(opt['v']), dict(step_no)
    g_out = {}
    for k, g in zip(keys, grads):
        g_out[k] = g
        if g is None:
            continue
        new_step[k] = step_no.get(k, 0) + 1
        new_sd[k], new_m[k], new_v[k] = adam_step(leaf[k].detach(), g, opt['m'][k].to(dtype), opt['v'][k].to(dtype), new_step[k], lr,
                                                  betas[0], betas[1])
    return dict(loss=loss.item(), recon=recon.detach(), outs=[None if o is None else o.detach() for o in outs], grads=g_out, sd=new_sd,
                opt={'m': new_m, 'v': new_v}, step=new_step)
