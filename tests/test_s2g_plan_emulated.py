"""CPU: the Speech2Gesture launch plan (tgb200/s2g_engine.py, train_eval/train_speech2gesture.py) executed on the NumPy restatement of the
C-ABI entry points (tests/cabi_emulator.py) and held to the GPU tests' own assertions against the reference-executed golden: checks the
host logic (im2col geometry and TensorFlow padding, the U-Net's skip wiring in the backward, the three discriminator passes, both Adam
bindings); the kernels themselves are checked on the B200 by tests/test_gpu_zzzz_speech2gesture.py."""
import pytest
import torch

import cabi_emulator
import test_gpu_zzzz_speech2gesture as GS

CPU = torch.device('cpu')


@pytest.fixture()
def emu():
    with cabi_emulator.installed() as e:
        yield e


def test_plumbing_kernels_restated(emu):
    GS.test_unet_plumbing_kernels_vs_torch(CPU)


@pytest.mark.parametrize('args', [(3, 16, 9, 8, 3, 3, 1, 1, True), (2, 17, 11, 4, 4, 4, 2, 2, True), (2, 14, 7, 8, 3, 3, 1, 1, False)])
def test_im2col_restated(emu, args):
    GS.test_im2col_col2im_vs_torch_conv(CPU, *args)


def test_eval_forward_plan_vs_reference_golden(emu):
    GS.test_eval_forward_vs_reference_golden(CPU)
    assert 'tg_im2col2d' in emu.calls and 'tg_resize_bilinear_fwd' in emu.calls


def test_train_iter_plan_vs_reference_golden(emu):
    GS.test_train_iter_vs_reference_golden(CPU)
    assert emu.calls.count('tg_adam_flat') == 4 and 'tg_col2im2d' in emu.calls
