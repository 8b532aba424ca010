"""tests/golden/forward_variants.npz: eval-mode PoseGenerator forwards of the UNMODIFIED reference module built with the other
constructor variants the drop-in boundary admits (args.input_context in audio / text / none, z_obj a Vocab / any truthy value / None;
scripts/model/multimodal_context_net.py:65-93,110-160).  tests/golden/train_variants.npz: one full reference train_iter_gan step
(epoch 11, dropout off) for each of those variants with the matching args.z_type - the step function branches on it
(scripts/train_eval/train_gan.py:58-86).  TEST INFRASTRUCTURE ONLY.   python -m oracle.make_golden_variants"""
import argparse
import os

import numpy as np
import torch

from . import synth
from . import trimodal_oracle as O
from .make_golden import OUT, MaskDrop, digest, golden_cfg, import_reference

VARIANTS = (('audio', 'speaker'), ('text', 'random'), ('none', None), ('both', 'random'), ('none', 'speaker'))


def main():
    ref_embed, ref_net, _, ref_vocab = import_reference()
    cfg = golden_cfg()
    B = 3
    inp = synth.make_inputs(cfg, B, seed=1)
    pre_seq = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    eps = synth.make_noise(cfg, B, seed=1).eps[0]
    store = {}
    for ctx, zm in VARIANTS:
        args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, input_context=ctx, hidden_size=cfg.hidden_size,
                                  n_layers=cfg.n_layers, dropout_prob=cfg.dropout_prob, freeze_wordembed=False)
        z_obj = None
        if zm == 'speaker':
            z_obj = ref_vocab.Vocab('vid', insert_default_tokens=False)
            while z_obj.n_words < cfg.n_speakers:
                z_obj.index_word(f'spk{z_obj.n_words}')
        elif zm == 'random':
            z_obj = 1                                                   # train.py:85
        G = ref_net.PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, None, z_obj=z_obj)
        G.load_state_dict(synth.with_tcn_aliases(synth.generator_state_dict_variant(cfg, ctx, zm)), strict=True)
        G.eval()
        orig_reparam, orig_randn = ref_embed.reparameterize, torch.randn
        ref_embed.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
        torch.randn = lambda *a, **k: eps.clone()
        try:
            with torch.no_grad():
                poses, z, mu, logvar = G(pre_seq, inp['in_text'], inp['in_audio'], inp['vid'] if zm == 'speaker' else None)
        finally:
            ref_embed.reparameterize, torch.randn = orig_reparam, orig_randn
        tag = f'{ctx}_{zm}'
        store[tag + '/poses'] = poses.numpy()
        if z is not None:
            store[tag + '/z'] = z.numpy()
        print(tag, float(poses.norm()))
    np.savez(os.path.join(OUT, 'forward_variants.npz'), **store)


def z_type_of(zm):
    return {'speaker': 'speaker', 'random': 'random', None: 'none'}[zm]


def main_train():
    ref_embed, ref_net, ref_gan, ref_vocab = import_reference()
    cfg = golden_cfg()
    B, epoch = 3, 11
    inp = synth.make_inputs(cfg, B, seed=1)
    noise = synth.golden_noise(cfg, B, 2, False)
    store = {}
    for ctx, zm in VARIANTS:
        args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, input_context=ctx, hidden_size=cfg.hidden_size,
                                  n_layers=cfg.n_layers, dropout_prob=cfg.dropout_prob, freeze_wordembed=False, z_type=z_type_of(zm),
                                  loss_warmup=cfg.loss_warmup, loss_gan_weight=cfg.loss_gan_weight, loss_regression_weight=cfg.loss_regression_weight,
                                  loss_kld_weight=cfg.loss_kld_weight, loss_reg_weight=cfg.loss_reg_weight)
        z_obj = None
        if zm == 'speaker':
            z_obj = ref_vocab.Vocab('vid', insert_default_tokens=False)
            while z_obj.n_words < cfg.n_speakers:
                z_obj.index_word(f'spk{z_obj.n_words}')
        elif zm == 'random':
            z_obj = 1
        G = ref_net.PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, None, z_obj=z_obj)
        D = ref_net.ConvDiscriminator(cfg.pose_dim)
        G.load_state_dict(synth.with_tcn_aliases(synth.generator_state_dict_variant(cfg, ctx, zm)), strict=True)
        D.load_state_dict(synth.discriminator_state_dict(cfg), strict=True)
        G.train(); D.train()
        G.gru.dropout = 0.0; D.gru.dropout = 0.0
        G.text_encoder.drop = MaskDrop()                                # empty queues: every nn.Dropout is the identity
        for blk in G.text_encoder.tcn.network:
            blk.net[3] = MaskDrop(); blk.net[7] = MaskDrop()
        eps_queue = [e.clone() for e in noise.eps]                      # three generator forwards after the warm-up (train_gan.py:30,50,67)
        n_fwd = 3 if zm is not None else 2                              # no third (divergence) forward without a z (:58)
        orig_reparam, orig_randn, orig_randperm = ref_embed.reparameterize, torch.randn, torch.randperm
        ref_embed.reparameterize = lambda mu, logvar: mu + eps_queue.pop(0) * torch.exp(0.5 * logvar)
        torch.randn = lambda *a, **k: eps_queue.pop(0)
        torch.randperm = lambda n, *a, **k: noise.perm.clone()
        try:
            g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
            d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
            ret = ref_gan.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'] if zm == 'speaker' else None,
                                         G, D, g_opt, d_opt)
        finally:
            ref_embed.reparameterize, torch.randn, torch.randperm = orig_reparam, orig_randn, orig_randperm
        assert len(eps_queue) == (3 - n_fwd if zm is not None else 3), (ctx, zm, len(eps_queue))
        tag = f'{ctx}_{zm}'
        for k, v in ret.items():
            store[f'{tag}/loss_{k}'] = np.float64(v)
        for k, p in G.named_parameters():
            if '.net.0.' in k or '.net.4.' in k:
                continue
            store[f'{tag}/ggrad/{k}'] = digest(p.grad if p.grad is not None else torch.zeros_like(p))[:34]
            store[f'{tag}/hasgrad/{k}'] = np.int64(p.grad is not None)
        print(tag, {k: float(v) for k, v in ret.items()})
    np.savez(os.path.join(OUT, 'train_variants.npz'), **store)


if __name__ == '__main__':
    main()
    main_train()
