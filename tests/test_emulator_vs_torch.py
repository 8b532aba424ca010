"""CPU: pins tests/cabi_emulator.py entry by entry - the GPU suite's own per-kernel tests (tests/test_gpu_kernels.py: every C-ABI
kernel family vs a float64 torch computation of the same op) are run here against the NumPy restatement.  The same assertions pass on
the B200 with the real kernels, so emulator and kernels agree with torch - and hence with each other - on every option these tests
exercise (gathered taps / stride / dilation / padding, BatchNorm prologue, GRU gate order and saved planes, losses, Adam, Philox)."""
import inspect

import pytest
import torch

import cabi_emulator
import test_gpu_kernels as GK

CPU = torch.device('cpu')


def _cases():
    out = []
    for name, fn in inspect.getmembers(GK, inspect.isfunction):
        if not name.startswith('test_'):
            continue
        params = [m for m in getattr(fn, 'pytestmark', []) if m.name == 'parametrize']
        cases = [c if isinstance(c, tuple) else (c,) for c in params[0].args[1]] if params else [()]
        for c in cases:
            if name == 'test_gru_layer_fwd_bwd' and c[0] > 128:
                continue                       # same code path as the smaller cases; python loops over 34 steps x 384 clips are slow
            out.append(pytest.param(fn, c, id='%s%s' % (name[5:], list(c) if c else '')))
    return out


@pytest.mark.parametrize('fn,case', _cases())
def test_kernel_family_on_emulator(fn, case):
    from tgb200 import config
    old = config.set_mode('fp32')
    try:
        with cabi_emulator.installed():
            fn(CPU, *case)
    finally:
        config.set_mode(old)
