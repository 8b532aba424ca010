"""GPU: the joint-embedding model (EmbeddingNet(mode='random') = ContextEncoder + PoseEncoderConv + PoseDecoderGRU,
train_iter_embed / eval_embed) through libtg_b200.so - the checks of tests/joint_checks.py, which the CPU suite runs on the emulated
launch plan.  fp32 mode is held to 1e-4, tf32 mode to 1e-2 on outputs and losses.  (Sorts last on purpose.)"""
import pytest
import torch

import joint_checks

# the cases below that have not run on hardware yet get a hard per-test limit: a hung persistent kernel must not hang the whole GPU run
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method='thread')]


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from tgb200 import _lib
    _lib.load()
    return torch.device('cuda:0')


@pytest.fixture()
def fp32():
    from tgb200 import config
    old = config.set_mode('fp32')
    yield
    config.set_mode(old)


@pytest.fixture()
def tf32():
    from tgb200 import config
    old = config.set_mode('tf32')
    yield
    config.set_mode(old)


def test_forwards_and_eval_embed_vs_reference_golden(dev, fp32):
    joint_checks.run_forwards(dev, tol=1e-4)


def test_train_iter_embed_speech_then_pose_vs_reference_golden(dev, fp32):
    joint_checks.run_two_steps(dev, tol=1e-4)


def test_fast_mode_forwards(dev, tf32):
    joint_checks.run_forwards(dev, tol=1e-2)


def test_evaluate_testset_joint_embedding_and_autoencoder(dev, fp32):
    joint_checks.run_evaluate_testset(dev)


def test_all_dropout_masks_vs_fp64_oracle(dev, fp32):
    joint_checks.run_batch_vs_fp64_oracle(dev, Bn=16, tol=1e-4)


def test_noise_is_drawn_on_the_device_and_cpu_tensors_are_refused(dev, tf32):
    from oracle import synth
    from tgb200 import _lib
    from train_eval.train_joint_embed import eval_embed, train_iter_embed
    cfg, args, net, opt = joint_checks.build(dev)
    net.train()
    data = {k: v.to(dev) for k, v in synth.make_inputs(cfg, 8, seed=90).items()}
    losses = [train_iter_embed(args, 0, data['in_text'], data['in_audio'], data['target'], net, opt, mode='random')['loss'] for _ in range(4)]
    assert all(l == l and l > 0 for l in losses) and len(set(losses)) == 4
    net.eval()
    a = eval_embed(data['in_text'], data['in_audio'], data['target'][:, :4], data['target'], net, mode='speech')[1]
    b = eval_embed(data['in_text'], data['in_audio'], data['target'][:, :4], data['target'], net, mode='speech')[1]
    assert bool(torch.isfinite(a).all()) and not torch.equal(a, b)            # ContextEncoder reparameterises in eval mode too (embedding_net.py:258)
    c = eval_embed(data['in_text'], data['in_audio'], data['target'][:, :4], data['target'], net, mode='pose')[1]
    d = eval_embed(data['in_text'], data['in_audio'], data['target'][:, :4], data['target'], net, mode='pose')[1]
    assert torch.equal(c, d)                                                 # the pose branch is deterministic
    with pytest.raises(_lib.TgError):
        train_iter_embed(args, 0, data['in_text'].cpu(), data['in_audio'].cpu(), data['target'].cpu(), net.train(), opt, mode='pose')


def test_fast_mode_batch32_vs_fp64_oracle(dev, tf32):
    joint_checks.run_batch_vs_fp64_oracle(dev, Bn=32, tol=1e-2, gtol=0.3)       # 0.09 predicted by the TF32-truncation model of the emulator


@pytest.mark.parametrize('ctx,zm', [('audio', 'speaker'), ('text', 'random'), ('none', None), ('both', 'random'), ('none', 'speaker')])
def test_train_iter_gan_constructor_variants_vs_reference_golden(dev, fp32, ctx, zm):
    """The generator's other constructor variants / args.z_type values through a full TRAINING step (the step function branches on
    z_type, train_gan.py:58-86) vs the reference's own run; added after this round's last GPU call - green on the emulated plan."""
    from tgb200 import config
    from test_gan_plan_emulated import run_train_variant
    old = config.set_graphs(False)
    try:
        run_train_variant(dev, ctx, zm)
    finally:
        config.set_graphs(old)
