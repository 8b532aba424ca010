"""Shared helpers of the GPU parity tests: build OUR modules with the deterministic synthetic weights of oracle/synth.py."""
import argparse

import torch

from oracle import synth
from oracle import trimodal_oracle as O


def make_args(cfg: O.HotPathConfig, dropout_prob=None):
    return argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, input_context='both', hidden_size=cfg.hidden_size,
                              n_layers=cfg.n_layers, dropout_prob=cfg.dropout_prob if dropout_prob is None else dropout_prob,
                              freeze_wordembed=False, z_type='speaker', loss_warmup=cfg.loss_warmup,
                              loss_gan_weight=cfg.loss_gan_weight, loss_regression_weight=cfg.loss_regression_weight,
                              loss_kld_weight=cfg.loss_kld_weight, loss_reg_weight=cfg.loss_reg_weight, wordembed_dim=cfg.wordembed_dim)


def build_ours(cfg: O.HotPathConfig, device, dropout_prob=None, seed=0):
    from model import vocab
    from model.multimodal_context_net import ConvDiscriminator, PoseGenerator
    args = make_args(cfg, dropout_prob)
    spk = vocab.Vocab('vid', insert_default_tokens=False)
    while spk.n_words < cfg.n_speakers:
        spk.index_word('spk%d' % spk.n_words)
    G = PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, None, z_obj=spk)
    D = ConvDiscriminator(cfg.pose_dim)
    gsd = synth.generator_state_dict(cfg, seed)
    dsd = synth.discriminator_state_dict(cfg, seed)
    G.load_state_dict(synth.with_tcn_aliases(gsd), strict=True)
    D.load_state_dict(dsd, strict=True)
    if device is not None:
        G, D = G.to(device), D.to(device)
    return args, G, D, gsd, dsd


def to_dev(d, device, dtype=None):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            v = v.to(device)
            if dtype is not None and v.is_floating_point():
                v = v.to(dtype)
        out[k] = v
    return out


def masks_to_ours(m, device):
    """oracle mask dict ([B,C,T] for tcn) -> channels-last [B*T, C] tensors on the device"""
    if m is None:
        return None
    out = {}
    for k, v in m.items():
        if k.startswith('tcn'):
            v = v.transpose(1, 2)
        out[k] = v.reshape(-1, v.shape[-1]).contiguous().to(device)
    return out


def masks_to_dev(m, device, dtype=None):
    if m is None:
        return None
    return {k: (v.to(device) if dtype is None else v.to(device=device, dtype=dtype)) for k, v in m.items()}
