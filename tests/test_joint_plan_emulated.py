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
