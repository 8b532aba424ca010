"""Generates tests/golden/seq2seq_step.npz by running the UNMODIFIED reference seq2seq modules (/root/reference/scripts,
build container only; never at test / bench time on the GPU box).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_seq2seq

The reference's nn.GRU inter-layer dropout cannot take an injected mask, so the golden step runs with dropout_prob = 0
(the same convention as oracle/make_golden.py); the mask path is checked oracle-vs-CUDA with injected masks instead."""
import argparse
import os

import numpy as np
import torch

from . import synth
from .make_golden import OUT, digest, import_reference
from .seq2seq_oracle import Seq2SeqConfig


def golden_cfg() -> Seq2SeqConfig:
    return Seq2SeqConfig(n_words=300)


def main():
    torch.set_num_threads(8)
    import_reference()
    import model.seq2seq_net as ref_s2s
    import train_eval.train_seq2seq as ref_train
    cfg = golden_cfg()
    B = 6
    args = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.0, n_pre_poses=cfg.n_pre_poses,
                              GAN_noise_size=0, loss_regression_weight=cfg.loss_regression_weight, loss_kld_weight=cfg.loss_kld_weight,
                              loss_reg_weight=cfg.loss_reg_weight)
    net = ref_s2s.Seq2SeqNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None)
    sd = synth.seq2seq_state_dict(cfg)
    net.load_state_dict(sd, strict=True)
    inp = synth.seq2seq_inputs(cfg, B, seed=3, max_len=9)
    store = {}
    net.eval()
    with torch.no_grad():
        out_eval = net(inp['in_text'], inp['lengths'], inp['target'], None)
    store['out_eval'] = out_eval.numpy()
    net.train()
    optim = torch.optim.Adam(net.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
    for it in range(2):
        # the reference returns only the loss; capture outputs and clipped grads through hooks on the instance
        ret = ref_train.train_iter_seq2seq(args, 0, inp['in_text'], inp['lengths'], inp['target'], net, optim)
        store[f'loss{it}'] = np.float64(ret['loss'])
        for k, p in net.named_parameters():
            store[f'grad{it}/' + k] = digest(p.grad)
        for k, v in net.state_dict().items():
            store[f'post{it}/' + k] = digest(v)
        print('step', it, ret)
    np.savez(os.path.join(OUT, 'seq2seq_step.npz'), **store)


if __name__ == '__main__':
    main()
