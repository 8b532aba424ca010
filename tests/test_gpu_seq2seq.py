"""GPU: the seq2seq baseline (config/seq2seq.yml) through the drop-in Seq2SeqNet / train_iter_seq2seq and
the C ABI, against (1) the golden step produced by the unmodified reference modules and (2) the float64 oracle at batch 128
with injected dropout masks."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import seq2seq_oracle as S
from oracle import synth
from oracle.make_golden import digest
from oracle.make_golden_seq2seq import golden_cfg
from test_oracle_golden import _digest_close, _post_close

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4          # north_star: fp32 mode within 1e-4 relative L2


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _args(cfg, dropout):
    return argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=dropout, n_pre_poses=cfg.n_pre_poses,
                              GAN_noise_size=0, loss_regression_weight=cfg.loss_regression_weight, loss_kld_weight=cfg.loss_kld_weight,
                              loss_reg_weight=cfg.loss_reg_weight)


def _build(cfg, dev, dropout=0.0):
    from model.seq2seq_net import Seq2SeqNet
    args = _args(cfg, dropout)
    net = Seq2SeqNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None)
    net.load_state_dict(synth.seq2seq_state_dict(cfg), strict=True)          # reference key names and shapes
    return args, net.to(dev)


def test_seq2seq_matches_reference_golden(dev):
    from tgb200 import config
    from train_eval.train_seq2seq import train_iter_seq2seq
    old = config.set_mode('fp32')
    try:
        cfg = golden_cfg()
        g = np.load(os.path.join(GOLDEN, 'seq2seq_step.npz'))
        args, net = _build(cfg, dev)
        inp = synth.seq2seq_inputs(cfg, 6, seed=3, max_len=9)
        text, target = inp['in_text'].to(dev), inp['target'].to(dev)
        net.eval()
        with torch.no_grad():
            out = net(text, inp['lengths'], target, None)
        assert rel_l2(out, g['out_eval']) < FP32_TOL, rel_l2(out, g['out_eval'])
        net.train()
        optim = torch.optim.Adam(net.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
        for it in range(2):
            ret = train_iter_seq2seq(args, 0, text, inp['lengths'], target, net, optim)
            ref = float(g[f'loss{it}'])
            assert abs(ret['loss'] - ref) <= FP32_TOL * abs(ref), (it, ret['loss'], ref)
            for k, p in net.named_parameters():
                _digest_close(digest(p.grad.cpu()), g[f'grad{it}/' + k], 1e-3)
            for k, v in net.state_dict().items():
                r = g[f'post{it}/' + k]
                if k.endswith('num_batches_tracked'):
                    assert int(v) == int(r[1])
                elif S.is_param(k):
                    _post_close(digest(v.cpu()), r, cfg.learning_rate, noisy=(it > 0 or k == 'decoder.decoder.pre_linear.0.bias'))
                else:
                    _digest_close(digest(v.cpu()), r, 1e-4)
    finally:
        config.set_mode(old)


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('tf32', 1e-2)])
def test_seq2seq_step_batch128_vs_fp64_oracle(dev, mode, tol):
    """Full training step at the benchmark shape with injected inter-layer dropout masks vs. the float64 oracle."""
    from tgb200 import config
    from train_eval import train_seq2seq as TS
    old = config.set_mode(mode)
    try:
        cfg = S.Seq2SeqConfig(n_words=2000)
        B, Tm = 128, 12
        args, net = _build(cfg, dev, dropout=cfg.dropout_prob)
        inp = synth.seq2seq_inputs(cfg, B, seed=5, max_len=Tm)
        rng = np.random.Generator(np.random.PCG64(99))
        H, T = cfg.hidden_size, cfg.n_poses
        keep = 1.0 - cfg.dropout_prob
        enc_mask = torch.from_numpy(((rng.random((B, Tm, 2 * H)) < keep) / keep).astype(np.float32))
        dec_mask = torch.from_numpy(((rng.random((T, B, H)) < keep) / keep).astype(np.float32))
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in synth.seq2seq_state_dict(cfg).items()}
        ref = S.train_iter_seq2seq_oracle(cfg, sd64, inp['in_text'], inp['lengths'], inp['target'].double(), None, step=1,
                                          enc_masks=[enc_mask.double()], dec_masks=[dec_mask[t].double() for t in range(1, T)])
        net.train()
        optim = torch.optim.Adam(net.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
        TS.inject_masks({'enc0': enc_mask.reshape(B * Tm, 2 * H).to(dev), 'dec0': dec_mask.to(dev)})
        ret = TS.train_iter_seq2seq(args, 0, inp['in_text'].to(dev), inp['lengths'], inp['target'].to(dev), net, optim)
        out = net.engine().ws['dec.outputs']
        assert rel_l2(out, ref['outputs']) < tol, rel_l2(out, ref['outputs'])
        assert abs(ret['loss'] - float(ref['loss'])) <= tol * abs(float(ref['loss']))
        worst = 0.0
        for k, p in net.named_parameters():
            r = ref['grads'][k]
            if float(r.norm()) < 1e-6:
                continue                                  # analytically zero (the Linear bias in front of the train-mode BatchNorm)
            worst = max(worst, rel_l2(p.grad, r))
        print('mode %s: outputs %.2e, worst gradient %.2e' % (mode, rel_l2(out, ref['outputs']), worst))
        assert worst < (10 * tol if mode == 'fp32' else 5 * tol), worst
    finally:
        config.set_mode(old)


def test_seq2seq_graph_replay_and_eval(dev):
    """The captured iteration keeps training (loss decreases over replays) and eval-mode forward runs at batch 128."""
    from train_eval.train_seq2seq import train_iter_seq2seq
    cfg = S.Seq2SeqConfig(n_words=2000)
    args, net = _build(cfg, dev, dropout=cfg.dropout_prob)
    inp = synth.seq2seq_inputs(cfg, 128, seed=6, max_len=10)
    text, target = inp['in_text'].to(dev), inp['target'].to(dev)
    net.train()
    optim = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))
    losses = [train_iter_seq2seq(args, 0, text, inp['lengths'], target, net, optim)['loss'] for _ in range(12)]
    assert any(s.graph is not None for s in net.engine()._graph_slots.values()), 'the iteration was never captured into a CUDA graph'
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert int(optim.state_dict()['state'][0]['step']) == 12
    net.eval()
    with torch.no_grad():
        out = net(text, inp['lengths'], target, None)
    assert out.shape == (128, cfg.n_poses, cfg.pose_dim) and bool(torch.isfinite(out).all())
    assert torch.equal(out[:, 0], target[:, 0])
