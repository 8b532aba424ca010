"""GPU: the validation loop around the path (scripts/train.py:234-329): device-side metrics kernel vs. the oracle /
reference golden, and evaluate_testset end to end for the seq2seq and multimodal_context generators."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import eval_oracle as E
from oracle import seq2seq_oracle as S
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def test_pose_metrics_kernel_matches_reference_golden(dev):
    from train_eval.evaluate import PoseMetrics
    g = np.load(os.path.join(GOLDEN, 'eval_metrics.npz'))
    m = PoseMetrics(dev)
    m.push(torch.from_numpy(g['out']).to(dev), torch.from_numpy(g['target']).to(dev), int(g['n_pre']))
    r = m.result()
    assert abs(r['loss'] - float(g['l1'])) <= 1e-6 * float(g['l1'])
    assert abs(r['joint_mae'] - float(g['mae'])) <= 1e-9 * float(g['mae'])
    assert abs(r['accel'] - float(g['accel'])) <= 1e-9 * float(g['accel'])


def test_evaluate_testset_seq2seq_matches_oracle_metrics(dev):
    from model.seq2seq_net import Seq2SeqNet
    from train_eval.evaluate import evaluate_testset
    cfg = S.Seq2SeqConfig(n_words=500)
    args = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.1, n_pre_poses=cfg.n_pre_poses, GAN_noise_size=0,
                              model='seq2seq', loss_regression_weight=250.0, loss_kld_weight=0.1, loss_reg_weight=25.0)
    net = Seq2SeqNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None)
    net.load_state_dict(synth.seq2seq_state_dict(cfg), strict=True)
    net = net.to(dev).train()
    batches, l1s, maes, accs, ns = [], [], [], [], []
    mean_dir_vec = np.zeros(27)
    for i, B in enumerate((16, 9)):
        inp = synth.seq2seq_inputs(cfg, B, seed=20 + i, max_len=8)
        batches.append((inp['in_text'], inp['lengths'], None, None, inp['target'], None, None, None))
        with torch.no_grad():
            out = S.seq2seq_forward({k: v for k, v in synth.seq2seq_state_dict(cfg).items()}, cfg, inp['in_text'], inp['lengths'], inp['target'], False)
        l1, mae, acc = E.batch_metrics(out.numpy(), inp['target'].numpy(), mean_dir_vec, cfg.n_pre_poses)
        l1s.append(l1); maes.append(mae); accs.append(acc); ns.append(B)
    from tgb200 import config
    old = config.set_mode('fp32')              # strict mode: this test pins the metric plumbing, not the TF32 tolerance
    try:
        ret = evaluate_testset(batches, net, None, None, args)
    finally:
        config.set_mode(old)
    w = np.array(ns, dtype=np.float64) / sum(ns)                     # AverageMeter: batch means weighted by batch size (train.py:289,306)
    assert net.training, 'evaluate_testset must restore the training flag (train.py:313)'
    assert abs(ret['loss'] - float((w * l1s).sum())) <= 2e-4 * float((w * l1s).sum())
    assert abs(ret['joint_mae'] - float((w * maes).sum())) <= 2e-4 * float((w * maes).sum())


def test_evaluate_testset_multimodal_with_fgd(dev):
    from gpu_util import build_ours
    from model.embedding_net import EmbeddingNet
    from model.embedding_space_evaluator import EmbeddingSpaceEvaluator
    from oracle import trimodal_oracle as O
    from train_eval.evaluate import evaluate_testset
    cfg = O.HotPathConfig(n_words=400, n_speakers=16)
    args, G, D, _, _ = build_ours(cfg, dev)
    args.model = 'multimodal_context'
    e_args = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.3, freeze_wordembed=False)
    enet = EmbeddingNet(e_args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None, 'pose')
    enet.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    ev = EmbeddingSpaceEvaluator.from_net(enet, cfg.n_pre_poses, dev)
    batches = []
    for i in range(3):
        inp = synth.make_inputs(cfg, 32, seed=40 + i)
        batches.append((None, None, inp['in_text'], None, inp['target'], inp['in_audio'], None, None))
    G.train()
    ret = evaluate_testset(batches, G, None, ev, args)
    assert set(ret) == {'loss', 'joint_mae', 'frechet', 'feat_dist'}
    assert all(np.isfinite(v) for v in ret.values()) and ret['loss'] > 0 and ret['joint_mae'] > 0
    assert G.training and ev.get_no_of_samples() == 3


def test_device_prefetcher_yields_identical_batches(dev):
    """Input staging: pinned and pageable host batches arrive on the device unchanged and in order."""
    from train_eval.staging import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    host = []
    for i in range(7):
        a = torch.randn(16, 1000, generator=g)
        t = torch.randint(0, 50, (16, 34), generator=g)
        host.append({'audio': a.pin_memory() if i % 2 == 0 else a, 'text': t, 'meta': ('clip%d' % i, [a[:2].clone()])})
    n = 0
    for i, b in enumerate(DevicePrefetcher(host, dev, depth=2)):
        assert b['audio'].is_cuda and b['text'].is_cuda and b['meta'][0] == 'clip%d' % i
        y = b['audio'] * 2.0                                  # consume on the current stream
        assert torch.equal(b['audio'].cpu(), host[i]['audio']) and torch.equal(b['text'].cpu(), host[i]['text'])
        assert torch.equal(b['meta'][1][0].cpu(), host[i]['meta'][1][0]) and torch.equal(y.cpu(), host[i]['audio'] * 2.0)
        n += 1
    assert n == 7
