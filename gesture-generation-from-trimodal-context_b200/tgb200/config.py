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


_WAV_FIRST = os.environ.get('TGB200_WAV_FIRST', '1') == '1'


def wav_first() -> bool:
    """Queue the WavEncoder forward before everything else of an iteration (default) instead of after the TextEncoderTCN chain.
    Measured on the graph-replayed batch-128 step: audio first 4.81 ms, text first 4.93 ms (profiles/r02_bench_wav_order.txt) - the audio
    chain's BatchNorm reductions are HBM-bound and overlap well with the tensor-core text chain when they start together."""
    return _WAV_FIRST


_DREAL_AT = os.environ.get('TGB200_DREAL_AT', 'gru1')


def d_real_at() -> str:
    """Where train_iter_gan forks the discriminator's pass over the REAL clips (independent of the generator): 'top' = at the start of
    the iteration, beside the audio / text encoders; 'concat' = when the GRU input exists; 'gru0' / 'gru1' = once the generator's first /
    second recurrent layer has been launched (the recurrence occupies 96 of 148 SMs for ~0.9 ms and nothing else is runnable then)."""
    return _DREAL_AT


_WAV_DGRAD_DIRECT = os.environ.get('TGB200_WAV_DGRAD_DIRECT', '1') == '1'


def wav_dgrad_direct() -> bool:
    """WavEncoder conv2-4 data gradients as accumulating-tap GEMMs (tg_conv_dgrad_tf32) instead of column GEMM + col2im."""
    return _WAV_DGRAD_DIRECT


_HEAD_PADDED = os.environ.get('TGB200_HEAD_PADDED', '1') == '1'


def head_padded() -> bool:
    """PoseGenerator output head: 152-float row pitch for the 150-wide hidden activation so its GEMMs run on the tensor cores (fast mode)."""
    return _HEAD_PADDED


_WGRAD_AFTER_DGRAD = os.environ.get('TGB200_WGRAD_AFTER_DGRAD', '1') == '1'


def wgrad_after_dgrad() -> bool:
    """GRU backward: fork a layer's weight-gradient GEMMs after (not before) its data-gradient GEMM (engine.GruPlan.backward)."""
    return _WGRAD_AFTER_DGRAD


_TCN_FUSED_ADD = os.environ.get('TGB200_TCN_FUSED_ADD', '1') == '1'


def tcn_fused_add() -> bool:
    """Fast mode: TemporalBlock's residual add + final ReLU ride on conv2's GEMM epilogue (y2 is not stored; backward: tg_tcn_res_bwd)."""
    return _TCN_FUSED_ADD


_FLAT_PRIO = os.environ.get('TGB200_FLAT_PRIO', '0') == '1'


def flat_prio() -> bool:
    """A/B switch: one high priority for the capture stream and every preparation stream instead of the graded ones (engine._Overlap)."""
    return _FLAT_PRIO


_NCCL_GRAPH = os.environ.get('TGB200_NCCL_GRAPH', '1') == '1'


def nccl_in_graph() -> bool:
    """Data parallel: capture the gradient all-reduces INSIDE the iteration's CUDA graph (one graph, no host-driven segment boundaries, the
    exchange of the recurrent layers' gradients overlapped with the rest of the backward).  TGB200_NCCL_GRAPH=0: graph segments split at
    the two collectives, NCCL launched eagerly in between (the round-1 scheme)."""
    return _NCCL_GRAPH


_D_FUSED = os.environ.get('TGB200_D_FUSED', '1') == '1'


def d_fused() -> bool:
    """ConvDiscriminator: run the 4-layer bidirectional GRU and both heads as one launch (csrc/dgru_stack.cu) instead of the per-layer
    plan (projection GEMM + recurrence + dropout multiply per layer, then sum + two head GEMMs)."""
    return _D_FUSED


def set_d_fused(v: bool) -> bool:
    global _D_FUSED
    old, _D_FUSED = _D_FUSED, bool(v)
    return old
