"""Arithmetic mode of the GEMM-class kernels (BASELINE.json north_star: an fp32 mode held to 1e-4 relative L2 and a
bf16/tf32 mode held to 1e-2).

  'fp32' : CUDA-core FFMA kernels everywhere (tg_conv_gemm_f32, tg_conv_wgrad_f32, tg_gru_layer_*): strict.
  'tf32' : the large GEMMs and the GRU recurrence run on the tcgen05 tensor cores with TF32 operands (10-bit mantissa)
           and fp32 accumulation in TMEM; small / oddly-shaped operators stay on the fp32 kernels.

Default 'tf32' (what stock PyTorch also does for cuDNN convolutions on Ampere+); override with TGB200_MODE=fp32 or
set_mode()."""
import os

_MODE = os.environ.get('TGB200_MODE', 'tf32')
assert _MODE in ('fp32', 'tf32'), _MODE


def mode() -> str:
    return _MODE


def set_mode(m: str) -> str:
    global _MODE
    assert m in ('fp32', 'tf32'), m
    old, _MODE = _MODE, m
    return old


def fast() -> bool:
    return _MODE == 'tf32'


_OVERLAP = os.environ.get('TGB200_OVERLAP', '1') == '1'


def overlap() -> bool:
    """Run off-critical-path launch sequences (weight gradients, the audio encoder, the discriminator's real-clip pass) on
    auxiliary CUDA streams so that they fill the SMs the latency-bound persistent kernels leave idle."""
    return _OVERLAP


def set_overlap(v: bool) -> bool:
    global _OVERLAP
    old, _OVERLAP = _OVERLAP, bool(v)
    return old


_GRAPHS = os.environ.get('TGB200_GRAPHS', '1') == '1'


def graphs() -> bool:
    """Capture the whole train_iter_gan launch sequence (~400 launches over 4 streams) into one CUDA graph after two eager
    iterations and replay it afterwards: removes the host-side launch cost, which otherwise exceeds the GPU time."""
    return _GRAPHS


def set_graphs(v: bool) -> bool:
    global _GRAPHS
    old, _GRAPHS = _GRAPHS, bool(v)
    return old
