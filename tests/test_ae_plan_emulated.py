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
