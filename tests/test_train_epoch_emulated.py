"""CPU: train.py::init_model / train_epoch (scripts/train.py:36-68,166-226) for the four model families, every step function's launch
plan executing on the NumPy C-ABI emulator: constructor switch, per-batch dispatch, speaker-id lookup, loss meters."""
import argparse

import numpy as np
import pytest
import torch

import cabi_emulator
from oracle import seq2seq_oracle as S
from oracle import synth
from oracle.make_golden import golden_cfg


class Lang:
    def __init__(self, n_words):
        self.n_words, self.word_embedding_weights = n_words, None


@pytest.fixture()
def emu():
    from tgb200 import config
    old_mode, old_graphs = config.set_mode('tf32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed() as e:
            yield e
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


def _loader(cfg, n_batches, B, speakers):
    out = []
    for i in range(n_batches):
        inp = synth.make_inputs(cfg, B, seed=200 + i)
        s2s = synth.seq2seq_inputs(S.Seq2SeqConfig(n_words=cfg.n_words), B, seed=300 + i, max_len=8)
        aux = {'vid': [speakers[int(v) % len(speakers)] for v in inp['vid']]}
        out.append((s2s['in_text'], s2s['lengths'], inp['in_text'], None, inp['target'], inp['in_audio'], torch.zeros(B, 4, 4), aux))
    return out


@pytest.mark.parametrize('model', ['multimodal_context', 'joint_embedding', 'gesture_autoencoder', 'seq2seq'])
def test_init_model_and_train_epoch(emu, model):
    import train
    from model import vocab
    cfg = golden_cfg()
    args = argparse.Namespace(model=model, n_poses=cfg.n_poses, n_pre_poses=cfg.n_pre_poses, wordembed_dim=cfg.wordembed_dim,
                              hidden_size=200 if model == 'seq2seq' else cfg.hidden_size, n_layers=2 if model == 'seq2seq' else cfg.n_layers,
                              dropout_prob=0.1 if model == 'seq2seq' else cfg.dropout_prob, freeze_wordembed=False, z_type='speaker',
                              input_context='both', loss_warmup=-1, loss_gan_weight=5.0, loss_regression_weight=500.0, loss_kld_weight=0.1,
                              loss_reg_weight=0.05, learning_rate=5e-4, discriminator_lr_weight=0.2, GAN_noise_size=0)
    spk = vocab.Vocab('vid', insert_default_tokens=False)
    names = ['spk%d' % i for i in range(cfg.n_speakers - 1)]
    for n in names:
        spk.index_word(n)
    torch.manual_seed(3)
    G, D, loss_fn = train.init_model(args, Lang(cfg.n_words), spk if model == 'multimodal_context' else None, cfg.pose_dim, torch.device('cpu'))
    assert (D is not None) == (model == 'multimodal_context') and (loss_fn is not None) == (model == 'seq2seq')
    g_opt = torch.optim.Adam(G.parameters(), lr=args.learning_rate, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=args.learning_rate * args.discriminator_lr_weight, betas=(0.5, 0.999)) if D is not None else None
    steps = []
    before = {k: v.clone() for k, v in G.state_dict().items()}
    ret = train.train_epoch(args, 0, _loader(cfg, 2, 3, names), G, D, g_opt, d_opt, speaker_model=spk if model == 'multimodal_context' else None,
                            on_step=lambda i, loss, n: steps.append((i, dict(loss), n)))
    assert [s[0] for s in steps] == [0, 1] and all(s[2] == 3 for s in steps)
    want = {'multimodal_context': {'loss', 'KLD', 'DIV_REG', 'gen', 'dis'}}.get(model, {'loss'})
    assert set(ret) == want and all(np.isfinite(v) for v in ret.values())
    assert abs(ret['loss'] - np.mean([s[1]['loss'] for s in steps])) < 1e-9 * abs(ret['loss'])
    assert any(not torch.equal(before[k], v) for k, v in G.state_dict().items() if v.is_floating_point())      # the optimiser moved the weights
    args.model = 'speech2gesture'                                  # train.py:59-62: the fifth family is constructed too
    g2, d2, l2 = train.init_model(args, Lang(cfg.n_words), None, cfg.pose_dim, torch.device('cpu'))
    assert type(g2).__name__ == 'Generator' and type(d2).__name__ == 'Discriminator' and isinstance(l2, torch.nn.L1Loss)
    with pytest.raises(NotImplementedError):
        args.model = 'no_such_model'
        train.init_model(args, Lang(cfg.n_words), None, cfg.pose_dim, torch.device('cpu'))
