"""CPU: pins the oracle (oracle/trimodal_oracle.py) against fixtures produced by the real
reference modules (oracle/make_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from oracle import trimodal_oracle as O
from oracle.make_golden import digest, golden_cfg
from conftest import GOLDEN, rel_l2

# Conv biases feeding a train-mode BatchNorm have an analytically ZERO gradient; what autograd
# produces is fp32 round-off, and Adam turns any non-zero round-off into a +-lr step.  Their
# post-step values are therefore noise in the reference itself and are excluded from parity.
ZERO_GRAD_KEYS = ('audio_encoder.feat_extractor.0.bias', 'audio_encoder.feat_extractor.3.bias',
                  'audio_encoder.feat_extractor.6.bias', 'pre_conv.0.bias', 'pre_conv.3.bias',
                  'pre_conv.1.bias', 'pre_conv.1.running_mean', 'pre_conv.4.running_mean')   # D: BN bias -> identity 'LeakyReLU(True)' -> conv -> train-mode BN
TOL = 2e-5          # fp32 CPU oracle vs fp32 CPU reference (different op order only)


def _digest_close(a, ref, tol):
    a = np.asarray(a); ref = np.asarray(ref)
    scale = max(abs(ref[0]), 1e-12)          # l2 norm of the tensor
    if scale < 1e-3:                          # analytically-zero grads (conv bias in front of BN): fp32 noise
        assert abs(a[0]) < 1e-3
        return
    assert abs(a[0] - ref[0]) <= tol * scale + 1e-9
    n = max(len(ref) - 2, 1)
    assert np.abs(a[2:] - ref[2:]).max() <= tol * 50 * scale / np.sqrt(n) + 5e-7


def _post_close(a, ref, lr, noisy=False):
    """Post-Adam weights.  On the first Adam step every element moves by ~lr*sign(g) whatever
    |g| is, so an element whose gradient is round-off noise may legitimately land 2*lr away;
    the bulk (median) must agree tightly.  Gradients themselves are checked tightly above."""
    d = np.abs(np.asarray(a)[2:] - np.asarray(ref)[2:])
    assert d.max() <= 2.2 * lr + 1e-6
    assert noisy or np.median(d) <= 2e-6 + 1e-5 * np.abs(np.asarray(ref)[2:]).max()


def test_forward_eval_matches_reference():
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'forward_eval.npz'))
    inp = synth.make_inputs(cfg, 3, seed=1)
    pre = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    eps = synth.make_noise(cfg, 3, seed=1).eps[0]
    gsd = synth.generator_state_dict(cfg)
    dsd = synth.discriminator_state_dict(cfg)
    with torch.no_grad():
        poses, z, mu, logvar = O.pose_generator_forward(gsd, cfg, pre, inp['in_text'], inp['in_audio'], inp['vid'], eps)
        audio = O.wav_encoder(gsd, 'audio_encoder', inp['in_audio'], False)
        text = O.text_encoder_tcn(gsd, 'text_encoder', inp['in_text'], cfg.n_layers)
        d_real = O.conv_discriminator_forward(dsd, cfg, inp['target'])
        d_fake = O.conv_discriminator_forward(dsd, cfg, poses)
    assert rel_l2(poses, g['poses']) < TOL
    assert rel_l2(audio, g['audio_feat']) < TOL
    assert rel_l2(text, g['text_feat']) < TOL
    assert rel_l2(z, g['z']) < TOL and rel_l2(mu, g['mu']) < TOL and rel_l2(logvar, g['logvar']) < TOL
    assert rel_l2(d_real, g['d_real']) < TOL and rel_l2(d_fake, g['d_fake']) < TOL


@pytest.mark.parametrize('tag,epoch,use_masks', [('train_e11', 11, True), ('train_e0', 0, False)])
def test_train_iter_matches_reference(tag, epoch, use_masks):
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, tag + '.npz'))
    inp = synth.make_inputs(cfg, 3, seed=1)
    noise = synth.golden_noise(cfg, 3, 2, use_masks)
    gsd = synth.generator_state_dict(cfg)
    dsd = synth.discriminator_state_dict(cfg)
    ret = O.train_iter_gan_oracle(cfg, epoch, gsd, dsd, synth.zeros_like_opt(gsd), synth.zeros_like_opt(dsd), 1,
                                  inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], noise)
    for k, v in ret['losses'].items():
        ref = float(g['loss_' + k])
        assert abs(v - ref) <= 1e-4 * abs(ref) + 1e-6, (k, v, ref)
    assert set('loss_' + k for k in ret['losses']) == set(k for k in g.files if k.startswith('loss_'))
    for k, gr in ret['g_grads'].items():
        _digest_close(digest(gr), g['ggrad/' + k], 2e-4)
    for k, v in ret['g_sd'].items():
        _post_close(digest(v), g['gpost/' + k], cfg.learning_rate, k in ZERO_GRAD_KEYS)
    for k, v in ret['d_sd'].items():
        _post_close(digest(v), g['dpost/' + k], cfg.learning_rate * cfg.discriminator_lr_weight, k in ZERO_GRAD_KEYS)


def test_embedding_net_and_fgd_match_reference():
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'embedding_fgd.npz'))
    esd = synth.embedding_net_state_dict(cfg)
    rng = np.random.Generator(np.random.PCG64(77))
    real = torch.from_numpy((0.5 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim))).astype(np.float32))
    fake = torch.from_numpy((1.0 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim)) + 0.3).astype(np.float32))
    with torch.no_grad():
        rf, rrec = O.embedding_net_pose_forward(esd, real)
        ff, frec = O.embedding_net_pose_forward(esd, fake)
    assert rel_l2(rf, g['real_feat']) < TOL and rel_l2(ff, g['fake_feat']) < TOL
    _digest_close(digest(rrec), g['real_recon'], 2e-5)
    fgd, fdist = O.fgd_scores(ff.numpy(), rf.numpy())
    assert abs(fgd - float(g['fgd'])) <= 1e-3 * abs(float(g['fgd']))
    assert abs(fdist - float(g['feat_dist'])) <= 1e-4 * abs(float(g['feat_dist']))


def test_seq2seq_oracle_matches_reference():
    """oracle/seq2seq_oracle.py vs. the reference Seq2SeqNet + train_iter_seq2seq (two consecutive steps, dropout 0):
    eval-mode outputs, loss, every clipped gradient, post-Adam weights and BatchNorm running statistics."""
    from oracle import seq2seq_oracle as S
    from oracle.make_golden_seq2seq import golden_cfg as s2s_cfg
    cfg = s2s_cfg()
    g = np.load(os.path.join(GOLDEN, 'seq2seq_step.npz'))
    sd = synth.seq2seq_state_dict(cfg)
    inp = synth.seq2seq_inputs(cfg, 6, seed=3, max_len=9)
    with torch.no_grad():
        out_eval = S.seq2seq_forward(sd, cfg, inp['in_text'], inp['lengths'], inp['target'], False)
    assert rel_l2(out_eval, g['out_eval']) < TOL
    opt = None
    for it in range(2):
        ret = S.train_iter_seq2seq_oracle(cfg, sd, inp['in_text'], inp['lengths'], inp['target'], opt, step=it + 1)
        ref = float(g[f'loss{it}'])
        assert abs(float(ret['loss']) - ref) <= 1e-4 * abs(ref), (it, float(ret['loss']), ref)
        for k, gr in ret['grads'].items():
            _digest_close(digest(gr), g[f'grad{it}/' + k], 5e-4)
        for k, v in ret['new_sd'].items():
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(g[f'post{it}/' + k][1])
            elif S.is_param(k):
                # the Linear bias in front of the train-mode BatchNorm has an analytically zero gradient (round-off only)
                _post_close(digest(v), g[f'post{it}/' + k], cfg.learning_rate, noisy=(it > 0 or k == 'decoder.decoder.pre_linear.0.bias'))
            else:
                _digest_close(digest(v), g[f'post{it}/' + k], 1e-4)
        sd, opt = ret['new_sd'], ret['opt']


def test_eval_metrics_oracle_matches_reference_function():
    """oracle/eval_oracle.py vs. the reference's own convert_dir_vec_to_pose + the metric lines of evaluate_testset (train.py:293-310)."""
    from oracle import eval_oracle as E
    g = np.load(os.path.join(GOLDEN, 'eval_metrics.npz'))
    l1, mae, acc = E.batch_metrics(g['out'], g['target'], g['mean_dir_vec'], int(g['n_pre']))
    assert abs(l1 - float(g['l1'])) < 1e-12 and abs(mae - float(g['mae'])) < 1e-12 and abs(acc - float(g['accel'])) < 1e-12
    pos = E.convert_dir_vec_to_pose(g['out'].astype(np.float64) + g['mean_dir_vec'])
    assert np.abs(pos - g['joint_poses']).max() < 1e-12


def test_forward_variants_oracle_matches_reference():
    """input_context / z_obj constructor variants (multimodal_context_net.py:72-93,139-153) vs. the reference module."""
    from oracle.make_golden_variants import VARIANTS
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'forward_variants.npz'))
    inp = synth.make_inputs(cfg, 3, seed=1)
    pre = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    eps = synth.make_noise(cfg, 3, seed=1).eps[0]
    for ctx, zm in VARIANTS:
        sd = synth.generator_state_dict_variant(cfg, ctx, zm)
        with torch.no_grad():
            poses, z, _, _ = O.pose_generator_forward(sd, cfg, pre, inp['in_text'], inp['in_audio'], inp['vid'], eps, input_context=ctx, z_mode=zm)
        assert rel_l2(poses, g[f'{ctx}_{zm}/poses']) < TOL, (ctx, zm)
        if zm is not None:
            assert rel_l2(z, g[f'{ctx}_{zm}/z']) < TOL


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_seq2seq_encoder_length_masks_equal_packed_sequence(seed):
    """The oracle restates pack_padded_sequence / pad_packed_sequence (seq2seq_net.py:52-56) with length masks; here it is checked
    directly against torch's own packed nn.GRU on random ragged batches (incl. length-1 rows and a full-length row)."""
    from oracle import seq2seq_oracle as S
    g = torch.Generator().manual_seed(seed)
    cfg = S.Seq2SeqConfig(n_words=50, wordembed_dim=12, hidden_size=10, n_layers=2)
    B, Tmax = 7, 9
    lengths = torch.sort(torch.randint(1, Tmax + 1, (B,), generator=g), descending=True).values
    lengths[0] = Tmax
    text = torch.zeros(B, Tmax, dtype=torch.int64)
    for b in range(B):
        text[b, :lengths[b]] = torch.randint(1, cfg.n_words, (int(lengths[b]),), generator=g)
    emb = torch.nn.Embedding(cfg.n_words, cfg.wordembed_dim)
    gru = torch.nn.GRU(cfg.wordembed_dim, cfg.hidden_size, cfg.n_layers, bidirectional=True)
    sd = {'encoder.embedding.weight': emb.weight.detach()}
    sd.update({'encoder.gru.' + k: v.detach() for k, v in gru.state_dict().items()})
    with torch.no_grad():
        packed = torch.nn.utils.rnn.pack_padded_sequence(emb(text.t()), lengths)
        out, hid = gru(packed, None)
        out, _ = torch.nn.utils.rnn.pad_packed_sequence(out)
        ref_out = (out[:, :, :cfg.hidden_size] + out[:, :, cfg.hidden_size:]).transpose(0, 1)
        enc, hidden = S.encoder_forward(sd, cfg, text, lengths)
    assert rel_l2(enc, ref_out) < 1e-6 and rel_l2(hidden, hid) < 1e-6


def test_eval_metrics_are_shift_invariant_and_linear():
    """Properties the device kernel relies on (tg_pose_eval_metrics drops the mean direction vector): the joint MAE and accel terms of
    train.py:293-310 do not depend on the vector added to both operands, and scale linearly with the error."""
    from oracle import eval_oracle as E
    rng = np.random.Generator(np.random.PCG64(5))
    tgt = rng.standard_normal((4, 34, 27)).astype(np.float32)
    err = (0.1 * rng.standard_normal((4, 34, 27))).astype(np.float32)
    m0 = E.batch_metrics(tgt + err, tgt, np.zeros(27), 4)
    m1 = E.batch_metrics(tgt + err, tgt, rng.standard_normal(27), 4)
    m2 = E.batch_metrics(tgt + 2 * err, tgt, np.zeros(27), 4)
    assert abs(m0[1] - m1[1]) < 1e-12 and abs(m0[2] - m1[2]) < 1e-12
    assert abs(m2[1] - 2 * m0[1]) < 1e-6 * m0[1] and abs(m2[2] - 2 * m0[2]) < 1e-6 * m0[2]
