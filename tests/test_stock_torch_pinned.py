"""oracle/stock_torch.py (the stock-PyTorch baseline bench.py times on the B200: nn.GRU / nn.Conv1d / weight_norm + autograd + torch.optim)
is pinned to the oracle, which is itself pinned to the executed reference (tests/test_oracle_golden.py).  CPU only."""
import dataclasses

import torch

from conftest import rel_l2
from oracle import stock_torch as ST
from oracle import synth
from oracle import trimodal_oracle as O
from test_oracle_golden import ZERO_GRAD_KEYS        # conv biases in front of a train-mode BatchNorm: analytically zero gradient


def _cfg():
    # dropout 0: nn.GRU's internal inter-layer dropout cannot take an injected mask (same restriction as the reference)
    return dataclasses.replace(O.HotPathConfig(n_words=300, n_speakers=12), dropout_prob=0.0, emb_dropout=0.0)


def test_stock_modules_eval_forward_matches_oracle():
    cfg = _cfg()
    gsd, dsd = synth.generator_state_dict(cfg), synth.discriminator_state_dict(cfg)
    G, D, _, _ = ST.build(cfg, gsd, dsd, 'cpu')
    G.eval(); D.eval()
    inp = synth.make_inputs(cfg, 3, seed=2)
    noise = synth.make_noise(cfg, 3, seed=3)
    pre = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    with torch.no_grad():
        out, z, mu, logvar = G(pre, inp['in_text'], inp['in_audio'], inp['vid'], noise.eps[0])
        ref = O.pose_generator_forward(gsd, cfg, pre, inp['in_text'], inp['in_audio'], inp['vid'], noise.eps[0], False, None, None)
        assert rel_l2(out, ref[0]) < 1e-5 and rel_l2(z, ref[1]) < 1e-6 and rel_l2(mu, ref[2]) < 1e-6 and rel_l2(logvar, ref[3]) < 1e-6
        assert rel_l2(D(inp['target']), O.conv_discriminator_forward(dsd, cfg, inp['target'], False, None, None)) < 1e-5


def test_stock_train_iter_matches_oracle_step():
    """One post-warm-up adversarial iteration: same losses and same post-Adam weights as the oracle (dropout off in D too: the stock
    ConvDiscriminator's GRU dropout is 0.3 like the reference's, so D runs in eval-statistics-free train mode only through p = 0)."""
    cfg = _cfg()
    gsd, dsd = synth.generator_state_dict(cfg), synth.discriminator_state_dict(cfg)
    G, D, g_opt, d_opt = ST.build(cfg, gsd, dsd, 'cpu')
    D.gru.dropout = 0.0
    B = 4
    inp = synth.make_inputs(cfg, B, seed=5)
    noise = synth.make_noise(cfg, B, seed=6, dropout=False)
    ret = ST.train_iter_gan_stock(cfg, 11, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt,
                                  eps=list(noise.eps), perm=noise.perm)
    ref = O.train_iter_gan_oracle(cfg, 11, gsd, dsd, synth.zeros_like_opt(gsd), synth.zeros_like_opt(dsd), 1,
                                  inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], noise)
    for k, v in ref['losses'].items():
        assert abs(ret[k] - v) <= 2e-4 * abs(v) + 1e-6, (k, ret[k], v)
    # the gradients the G update used stay in .grad after the step (comparing post-Adam weights would compare sign(g) of noise-level gradients)
    grads = {k.replace('text_encoder.drop', 'text_encoder.dropout'): p.grad for k, p in G.named_parameters()}
    for k, v in ref['g_grads'].items():
        if k in ZERO_GRAD_KEYS or v.abs().max() < 1e-7:
            continue
        assert rel_l2(grads[k], v) < 2e-3, (k, rel_l2(grads[k], v))
    sd = G.state_dict()
    for k, v in ref['g_sd'].items():
        if k.endswith('num_batches_tracked'):
            assert int(sd[k]) == int(v), k
        elif k.endswith(('running_mean', 'running_var')):
            assert rel_l2(sd[k], v) < 1e-4, k
