"""Generates tests/golden/ae_train.npz by EXECUTING the reference's own auto-encoder trainer:

  * `train_iter` of scripts/train_feature_extractor.py:54-97 is extracted from the source file with `ast` (importing the module
    needs matplotlib / lmdb / the Human3.6M loader) and executed unmodified, two consecutive steps on the reference
    EmbeddingNet(mode='pose') with a torch.optim.Adam built as at train_feature_extractor.py:134;
  * `train_iter_embed` and `eval_embed` of scripts/train_eval/train_joint_embed.py are imported and executed as they are
    (one step from the same initial weights; eval on the initial weights).

Weights are oracle/synth.py's deterministic EmbeddingNet state dict loaded with strict=True.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_ae"""
import argparse
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import synth
from .make_golden import OUT, digest, golden_cfg, import_reference

SRC = '/root/reference/scripts/train_feature_extractor.py'
LR = 5e-4
B = 6


def reference_train_iter():
    tree = ast.parse(open(SRC).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'train_iter']
    assert len(keep) == 1
    ns = {'torch': torch, 'F': F}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, 'exec'), ns)
    return ns['train_iter']


def targets(cfg, n, seed):
    return synth.make_inputs(cfg, n, seed=seed)['target']


def build(ref_embed, cfg):
    args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, loss_kld_weight=0.1, loss_regression_weight=500.0)
    net = ref_embed.EmbeddingNet(args, cfg.pose_dim, cfg.n_poses, None, None, None, mode='pose')       # train_feature_extractor.py:133
    net.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    opt = torch.optim.Adam(net.parameters(), lr=LR, betas=(0.5, 0.999))                                 # :134
    return args, net, opt


def main():
    torch.set_num_threads(8)
    ref_embed, _, _, _ = import_reference()
    import train_eval.train_joint_embed as ref_joint
    cfg = golden_cfg()
    train_iter = reference_train_iter()
    store = {'lr': np.float64(LR)}

    # ---- train_feature_extractor.train_iter, two consecutive steps
    args, net, opt = build(ref_embed, cfg)
    net.train()
    for step in (1, 2):
        tgt = targets(cfg, B, 20 + step)
        store[f'fx{step}/target'] = tgt.numpy()
        ret = train_iter(args, 0, tgt, net, opt)
        store[f'fx{step}/loss'] = np.float64(ret['loss'])
        for k, p in net.named_parameters():
            store[f'fx{step}/grad/{k}'] = digest(p.grad if p.grad is not None else torch.zeros_like(p))
        for k, v in net.state_dict().items():
            store[f'fx{step}/post/{k}'] = digest(v)
        print('train_feature_extractor.train_iter step', step, ret)

    # ---- train_joint_embed.train_iter_embed on the pose auto-encoder (mode=None -> net.mode == 'pose'), one step
    args, net, opt = build(ref_embed, cfg)
    net.train()
    tgt = targets(cfg, B, 21)
    ret = ref_joint.train_iter_embed(args, 0, None, None, tgt, net, opt)
    store['je/loss'] = np.float64(ret['loss'])
    for k, p in net.named_parameters():
        store[f'je/grad/{k}'] = digest(p.grad if p.grad is not None else torch.zeros_like(p))
    for k, v in net.state_dict().items():
        store[f'je/post/{k}'] = digest(v)
    print('train_joint_embed.train_iter_embed', ret)

    # ---- train-mode forward (returned tuple) and eval_embed on the initial weights
    args, net, opt = build(ref_embed, cfg)
    net.train()
    with torch.no_grad():
        _, _, _, feat, mu, logvar, recon = net(None, None, None, tgt, None, variational_encoding=False)
    store['fwd_train/feat'] = feat.numpy(); store['fwd_train/logvar'] = logvar.numpy(); store['fwd_train/recon'] = recon.numpy()
    for k, v in net.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            store[f'fwd_train/post/{k}'] = digest(v)
    args, net, opt = build(ref_embed, cfg)
    net.eval()
    with torch.no_grad():
        loss, recon = ref_joint.eval_embed(None, None, None, tgt, net)
    store['eval/loss'] = np.float64(loss.item()); store['eval/recon'] = recon.numpy()
    print('eval_embed', loss.item())
    np.savez(os.path.join(OUT, 'ae_train.npz'), **store)


if __name__ == '__main__':
    main()
