"""CPU: the seq2seq training plan (tgb200/seq2seq_engine.py, train_eval/train_seq2seq.py) and the device-side validation loop
(train_eval/evaluate.py) executed on the NumPy restatement of the C-ABI entries they call (tests/cabi_emulator.py), held to the GPU
tests' own assertions (tests/test_gpu_seq2seq.py, tests/test_gpu_evaluate.py) against the reference-executed goldens."""
import pytest
import torch

import cabi_emulator
import test_gpu_evaluate as GE
import test_gpu_seq2seq as GS

CPU = torch.device('cpu')


@pytest.fixture()
def emu_fp32():
    from tgb200 import config
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed() as e:
            yield e
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


def test_seq2seq_plan_vs_reference_golden(emu_fp32):
    """Eval forward + two consecutive train_iter_seq2seq steps (packed bi-GRU encoder, 33-step attention decoder, custom_loss,
    clip_grad_norm_, Adam) vs the reference's own run."""
    GS.test_seq2seq_matches_reference_golden(CPU)
    assert emu_fp32.calls.count('tg_clip_scale') == 2 and emu_fp32.calls.count('tg_attn_bwd') == 2 * 33


def test_pose_metrics_plan_vs_reference_golden(emu_fp32):
    GE.test_pose_metrics_kernel_matches_reference_golden(CPU)


def test_evaluate_testset_seq2seq_plan(emu_fp32):
    GE.test_evaluate_testset_seq2seq_matches_oracle_metrics(CPU)


def test_evaluate_testset_multimodal_with_fgd_plan(emu_fp32):
    GE.test_evaluate_testset_multimodal_with_fgd(CPU)


def test_noise_drawing_plans(emu_fp32):
    """Plans that draw their own noise (Philox counters advanced on the 'device'): RNG statistics, then a training step and an
    eval-mode forward without injected noise - finite, and a fresh draw on every call."""
    import numpy as np
    import test_gpu_kernels as GK
    from gpu_util import build_ours
    from oracle import synth
    from oracle import trimodal_oracle as O
    from train_eval.train_gan import train_iter_gan
    GK.test_philox_rng_statistics(CPU)
    cfg = O.HotPathConfig(n_words=300, n_speakers=12)
    args, G, D, _, _ = build_ours(cfg, None)
    G.train(); D.train()
    inp = synth.make_inputs(cfg, 2, seed=1)
    go = torch.optim.Adam(G.parameters(), lr=5e-4, betas=(0.5, 0.999)); do = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
    rets = [train_iter_gan(args, 11, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, go, do) for _ in range(2)]
    assert all(np.isfinite(v) for r in rets for v in r.values()) and set(rets[0]) == {'loss', 'KLD', 'DIV_REG', 'gen', 'dis'}
    assert rets[0]['loss'] != rets[1]['loss']
    assert emu_fp32.calls.count('tg_philox_randperm') == 3 and 'tg_philox_dropout_mask' in emu_fp32.calls
    G.eval()
    pre = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    with torch.no_grad():
        a = G(pre, inp['in_text'], inp['in_audio'], inp['vid'])[0].clone()
        b = G(pre, inp['in_text'], inp['in_audio'], inp['vid'])[0].clone()
    assert bool(torch.isfinite(a).all()) and not torch.equal(a, b)          # reparameterize draws noise in eval mode too


@pytest.mark.parametrize('mode', ['fp32', 'tf32'])
def test_seq2seq_batch128_plan_vs_fp64_oracle(mode):
    """The benchmark shape (batch 128, text length 12) with injected inter-layer dropout masks, both arithmetic modes' plans."""
    from tgb200 import config
    old = config.set_graphs(False)
    try:
        with cabi_emulator.installed():
            GS.test_seq2seq_step_batch128_vs_fp64_oracle(CPU, mode, 1e-4)      # the emulator does not round to TF32: fp32-class agreement in both
    finally:
        config.set_graphs(old)


@pytest.mark.parametrize('kw', [dict(hidden_size=64, n_layers=1), dict(hidden_size=96, n_layers=3), dict(n_pre_poses=2)],
                         ids=lambda kw: ','.join('%s=%s' % it for it in kw.items()))
def test_seq2seq_other_hyper_parameters(emu_fp32, kw):
    """The seq2seq plan is not hard-wired to config/seq2seq.yml: other layer counts, hidden widths and seed-pose counts vs the fp64 oracle."""
    import argparse
    from conftest import rel_l2
    from model.seq2seq_net import Seq2SeqNet
    from oracle import seq2seq_oracle as S
    from oracle import synth
    from train_eval.train_seq2seq import train_iter_seq2seq
    cfg = S.Seq2SeqConfig(n_words=150, dropout_prob=0.0, **kw)
    args = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.0, n_pre_poses=cfg.n_pre_poses, GAN_noise_size=0,
                              loss_regression_weight=250.0, loss_kld_weight=0.1, loss_reg_weight=25.0)
    net = Seq2SeqNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None)
    sd = synth.seq2seq_state_dict(cfg)
    net.load_state_dict(sd, strict=True)
    net.train()
    inp = synth.seq2seq_inputs(cfg, 5, seed=2, max_len=7)
    ref = S.train_iter_seq2seq_oracle(cfg, {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, inp['in_text'], inp['lengths'],
                                      inp['target'].double(), None, step=1)
    ret = train_iter_seq2seq(args, 0, inp['in_text'], inp['lengths'], inp['target'], net, torch.optim.Adam(net.parameters(), lr=cfg.learning_rate))
    assert abs(ret['loss'] - float(ref['loss'])) <= 1e-5 * abs(float(ref['loss']))
    for k, p in net.named_parameters():
        if float(ref['grads'][k].norm()) > 1e-6:
            assert rel_l2(p.grad, ref['grads'][k]) < 1e-4, k
