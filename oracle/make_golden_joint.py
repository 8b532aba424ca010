"""Generates tests/golden/joint_embed.npz by executing the reference's EmbeddingNet(mode='random') (ContextEncoder + PoseEncoderConv +
PoseDecoderGRU, scripts/model/embedding_net.py) and train_iter_embed / eval_embed (scripts/train_eval/train_joint_embed.py):
  * eval-mode and train-mode forwards for input_mode 'speech' and 'pose',
  * two consecutive train_iter_embed steps, the first on the 'speech' branch, the second on the 'pose' branch (mode='random' with
    random.random patched to a queue: embedding_net.py:295-296), so that the second step exercises Adam's skipping of the
    parameters that got no gradient in either step.
Noise seams: model.embedding_net.reparameterize (eps), nn.Dropout instances swapped for mask queues, GRU inter-layer dropout p=0
(cannot take a mask).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_joint"""
import argparse
import os
import random

import numpy as np
import torch

from . import synth
from .make_golden import OUT, MaskDrop, digest, golden_cfg, import_reference

LR = 5e-4
B = 4


def build(ref_embed, cfg):
    args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, hidden_size=cfg.hidden_size, n_layers=cfg.n_layers,
                              dropout_prob=cfg.dropout_prob, freeze_wordembed=False, loss_kld_weight=0.1, loss_regression_weight=500.0)
    net = ref_embed.EmbeddingNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None, mode='random')   # train.py:60-62
    net.load_state_dict(synth.with_tcn_aliases(synth.joint_embedding_state_dict(cfg)), strict=True)
    net.decoder.gru.dropout = 0.0
    return args, net


def install_mask_queues(net):
    te = net.context_encoder.text_encoder
    drops = {'emb': MaskDrop()}
    te.drop = drops['emb']
    for i, blk in enumerate(te.tcn.network):
        for j, pos in ((1, 3), (2, 7)):
            md = MaskDrop()
            drops[f'tcn{i}_{j}'] = md
            blk.net[pos] = md
    return drops


def main():
    torch.set_num_threads(8)
    ref_embed, _, _, _ = import_reference()
    import train_eval.train_joint_embed as ref_joint
    cfg = golden_cfg()
    inp = synth.make_inputs(cfg, B, seed=61)
    store = {'lr': np.float64(LR)}
    orig = ref_embed.reparameterize
    try:
        # ---- forwards
        for training in (False, True):
            for mode in ('speech', 'pose'):
                args, net = build(ref_embed, cfg)
                net.train(training)
                eps = synth.make_noise(cfg, B, seed=62).eps[0].repeat(1, 2)[:, :32].contiguous()
                ref_embed.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
                if training:
                    install_mask_queues(net)                      # empty queues: dropout off
                with torch.no_grad():
                    outs = net(inp['in_text'], inp['in_audio'], inp['target'][:, :cfg.n_pre_poses], inp['target'], mode, variational_encoding=False)
                tag = f"fwd_{'train' if training else 'eval'}_{mode}"
                for name, o in zip(('c_feat', 'c_mu', 'c_lv', 'p_feat', 'p_mu', 'p_lv', 'out'), outs):
                    store[f'{tag}/{name}'] = o.numpy()
                store['eps'] = eps.numpy()
        # ---- eval_embed, mode='speech' (train.py:270)
        args, net = build(ref_embed, cfg)
        net.eval()
        with torch.no_grad():
            loss, recon = ref_joint.eval_embed(inp['in_text'], inp['in_audio'], inp['target'][:, :cfg.n_pre_poses], inp['target'], net, mode='speech')
        store['eval/loss'] = np.float64(loss.item()); store['eval/recon'] = recon.numpy()
        # ---- two train_iter_embed steps with mode='random': coin -> 'speech', then 'pose'
        args, net = build(ref_embed, cfg)
        net.train()
        drops = install_mask_queues(net)
        opt = torch.optim.Adam(net.parameters(), lr=LR, betas=(0.5, 0.999))
        coins = [0.9, 0.1]                                        # > 0.5 -> 'speech' (embedding_net.py:296)
        orig_random = random.random
        random.random = lambda: coins.pop(0)
        try:
            for step in (1, 2):
                noise = synth.golden_noise(cfg, B, 70 + step, True)
                for k, md in drops.items():
                    md.queue.append(noise.g_masks[0][k])
                e = noise.eps[0].repeat(1, 2)[:, :32].contiguous()
                ref_embed.reparameterize = lambda mu, logvar, e=e: mu + e * torch.exp(0.5 * logvar)
                data = synth.make_inputs(cfg, B, seed=63 + step)
                ret = ref_joint.train_iter_embed(args, 0, data['in_text'], data['in_audio'], data['target'], net, opt, mode='random')
                store[f'step{step}/loss'] = np.float64(ret['loss'])
                for k, p in net.named_parameters():
                    store[f'step{step}/grad/{k}'] = digest(p.grad if p.grad is not None else torch.zeros_like(p))
                    store[f'step{step}/hasgrad/{k}'] = np.int64(p.grad is not None)
                for k, v in net.state_dict().items():
                    if '.tcn.network.' in k and ('.net.0.' in k or '.net.4.' in k):
                        continue                                  # TemporalBlock alias keys (tcn.py:30-31)
                    store[f'step{step}/post/{k}'] = digest(v)
                print('train_iter_embed step', step, ret)
        finally:
            random.random = orig_random
        assert not coins
    finally:
        ref_embed.reparameterize = orig
    np.savez(os.path.join(OUT, 'joint_embed.npz'), **store)
    print('eval_embed', float(store['eval/loss']))


if __name__ == '__main__':
    main()
