"""CPU: pins tests/cabi_emulator.py entry by entry - the GPU suite's own per-kernel tests (tests/test_gpu_kernels.py: every C-ABI
kernel family vs a float64 torch computation of the same op) are run here against the NumPy restatement.  The same assertions pass on
the B200 with the real kernels, so emulator and kernels agree with torch - and hence with each other - on every option these tests
exercise (gathered taps / stride / dilation / padding, BatchNorm prologue, GRU gate order and saved planes, losses, Adam, Philox)."""
import inspect
import itertools

import pytest
import torch

import cabi_emulator
import test_gpu_kernels as GK
import test_gpu_tf32 as GT

CPU = torch.device('cpu')


def _cases(mod):
    out = []
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith('test_'):
            continue
        # stacked @parametrize decorators: pytestmark lists the innermost first; bind by argument name
        marks = [m for m in getattr(fn, 'pytestmark', []) if m.name == 'parametrize']
        axes = []
        for m in marks:
            names = [a.strip() for a in m.args[0].split(',')]
            axes.append([dict(zip(names, c if isinstance(c, tuple) else (c,))) for c in m.args[1]])
        for combo in itertools.product(*axes):
            kw = {k: v for d in combo for k, v in d.items()}
            if kw.get('B', 0) * kw.get('T', 0) > 128 * 34:
                continue                       # same code path as the smaller cases; python loops over the time steps are slow
            out.append(pytest.param(fn, kw, id='%s%s' % (name[5:], list(kw.values()) if kw else '')))
    return out


@pytest.mark.parametrize('fn,kw', _cases(GK) + _cases(GT))
def test_kernel_family_on_emulator(fn, kw):
    """test_gpu_kernels.py: the fp32 families; test_gpu_tf32.py: the tensor-core entries (restated without the TF32 rounding, so the
    1e-2-class tolerances of those tests are met with a wide margin - what is pinned is the operand / stride / epilogue contract)."""
    from tgb200 import config
    old = config.set_mode('fp32')
    try:
        with cabi_emulator.installed():
            fn(CPU, **kw)
    finally:
        config.set_mode(old)
