"""CPU: the auto-encoder trainer's LAUNCH PLAN (tgb200/embed_engine.py, train_eval/train_joint_embed.py, train_feature_extractor.py)
executed on the NumPy restatement of the C-ABI entry points it calls (tests/cabi_emulator.py) and compared with the reference-executed
golden and the fp64 oracle.  Checks the host logic - operand strides, prologues, the transposed-convolution mapping, the BatchNorm
backward chain, the Adam binding; the kernels themselves are checked on the GPU by tests/test_gpu_zz_autoencoder.py, which runs the
same checks (tests/ae_checks.py) through libtg_b200.so."""
import os

import numpy as np
import pytest
import torch

import ae_checks
import cabi_emulator
from conftest import GOLDEN, rel_l2
from oracle import synth
from oracle.make_golden import golden_cfg

CPU = torch.device('cpu')


@pytest.fixture()
def emu():
    with cabi_emulator.installed() as e:
        yield e


def test_emulator_reproduces_the_gpu_proven_eval_forward(emu):
    """Pins the emulator itself: the eval-mode EmbeddingNet plan (tgb200.engine.EmbeddingEngine), which is parity-green on the B200
    against the same fixture (test_gpu_parity.py::test_embedding_net_and_fgd_vs_reference_golden), must reproduce it here too."""
    from model.embedding_net import EmbeddingNet
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'embedding_fgd.npz'))
    E = EmbeddingNet(None, cfg.pose_dim, cfg.n_poses, None, None, None, 'pose')
    E.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    E.eval()
    rng = np.random.Generator(np.random.PCG64(77))
    real = torch.from_numpy((0.5 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim))).astype(np.float32))
    out = E(None, None, None, real[:64], 'pose', variational_encoding=False)
    assert rel_l2(out[3], g['real_feat'][:64]) < 2e-5
    assert 'tg_conv_gemm_f32' in emu.calls


def test_feature_extractor_train_iter_two_steps(emu):
    ae_checks.run_feature_extractor_two_steps(CPU)
    assert emu.calls.count('tg_adam_flat') == 2 and emu.calls.count('tg_ae_recon_loss') == 2


def test_train_iter_embed_pose_mode(emu):
    ae_checks.run_train_iter_embed(CPU)


def test_train_forward_and_eval_embed(emu):
    ae_checks.run_forward_and_eval(CPU)


def test_three_steps_vs_fp64_oracle(emu):
    ae_checks.run_full_batch_vs_fp64_oracle(CPU, B=16, steps=3)


def test_emulator_is_uninstalled_afterwards():
    from tgb200 import _lib
    assert not isinstance(_lib._lib, cabi_emulator.EmuLib)
    assert _lib.TRACE_ONLY == (os.environ.get('TGB200_TRACE_ONLY', '') == '1')


def test_resume_from_saved_model_and_optimizer_state(emu):
    """Checkpoint / resume: the caller's torch.optim.Adam is bound to the flat arena, so optim.state_dict() holds the real moments and step
    count; a fresh net + optimiser that load_state_dict() both continue bit-for-bit where the original continues."""
    import copy
    import train_feature_extractor as tfx
    cfg, net, opt = ae_checks.build(CPU)
    net.train()
    batches = [synth.make_inputs(cfg, 6, seed=70 + i)['target'] for i in range(3)]
    for b in batches[:2]:
        tfx.train_iter(None, 0, b, net, opt)
    saved_model, saved_opt = copy.deepcopy(net.state_dict()), copy.deepcopy(opt.state_dict())
    assert int(saved_opt['state'][0]['step']) == 2 and float(saved_opt['state'][0]['exp_avg'].abs().sum()) > 0
    tfx.train_iter(None, 0, batches[2], net, opt)                       # the original run goes on
    cfg, net2, opt2 = ae_checks.build(CPU)
    net2.train()
    tfx.train_iter(None, 0, batches[0], net2, opt2)                     # the resumed objects have already been used (arena bound) ...
    net2.load_state_dict(saved_model)
    opt2.load_state_dict(saved_opt)                                     # ... then a checkpoint is loaded into them
    tfx.train_iter(None, 0, batches[2], net2, opt2)
    for (k, a), b in zip(net.state_dict().items(), net2.state_dict().values()):
        assert torch.equal(a, b), k
    assert int(opt2.state_dict()['state'][0]['step']) == 3
