"""Deterministic synthetic weights and inputs shared by the oracle, the golden-vector
generator, the GPU parity tests and bench.py.  TEST / MEASUREMENT INFRASTRUCTURE.

Weights are drawn from numpy's PCG64 stream (stable across numpy versions) so the very
same tensors can be rebuilt on the GPU box without shipping 50 MB fixtures.  Key names
and shapes are the reference's state_dict (checked against the real reference modules by
oracle/make_golden.py).  Input distributions follow SURVEY.md section 8d.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np
import torch

from .trimodal_oracle import HotPathConfig


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _bn(sd, rng, name, c):
    sd[name + '.weight'] = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32))
    sd[name + '.bias'] = torch.from_numpy((0.1 * rng.standard_normal(c)).astype(np.float32))
    sd[name + '.running_mean'] = torch.from_numpy((0.05 * rng.standard_normal(c)).astype(np.float32))
    sd[name + '.running_var'] = torch.from_numpy((1.0 + 0.2 * rng.random(c)).astype(np.float32))
    sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)


def _lin(sd, rng, name, shape, fan_in, bias=True, scale=1.0):
    s = scale / np.sqrt(fan_in)
    sd[name + '.weight'] = torch.from_numpy((s * rng.standard_normal(shape)).astype(np.float32))
    if bias:
        sd[name + '.bias'] = torch.from_numpy((s * rng.standard_normal(shape[0])).astype(np.float32))


def _gru(sd, rng, prefix, in_size, hidden, layers):
    s = 1.0 / np.sqrt(hidden)
    for l in range(layers):
        for suf in ('', '_reverse'):
            isz = in_size if l == 0 else 2 * hidden
            sd[f'{prefix}.weight_ih_l{l}{suf}'] = torch.from_numpy((s * rng.standard_normal((3 * hidden, isz))).astype(np.float32))
            sd[f'{prefix}.weight_hh_l{l}{suf}'] = torch.from_numpy((s * rng.standard_normal((3 * hidden, hidden))).astype(np.float32))
            sd[f'{prefix}.bias_ih_l{l}{suf}'] = torch.from_numpy((s * rng.standard_normal(3 * hidden)).astype(np.float32))
            sd[f'{prefix}.bias_hh_l{l}{suf}'] = torch.from_numpy((s * rng.standard_normal(3 * hidden)).astype(np.float32))


def generator_state_dict(cfg: HotPathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict of PoseGenerator (multimodal_context_net.py:64-104) in reference key order."""
    rng = _rng(1000 + seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    p = 'audio_encoder.feat_extractor'
    _lin(sd, rng, p + '.0', (16, 1, 15), 15)
    _bn(sd, rng, p + '.1', 16)
    _lin(sd, rng, p + '.3', (32, 16, 15), 16 * 15)
    _bn(sd, rng, p + '.4', 32)
    _lin(sd, rng, p + '.6', (64, 32, 15), 32 * 15)
    _bn(sd, rng, p + '.7', 64)
    _lin(sd, rng, p + '.9', (32, 64, 15), 64 * 15)
    E, Hh = cfg.wordembed_dim, cfg.hidden_size
    sd['text_encoder.embedding.weight'] = torch.from_numpy(rng.standard_normal((cfg.n_words, E)).astype(np.float32))
    for i in range(cfg.n_layers):
        cin = E if i == 0 else Hh
        q = f'text_encoder.tcn.network.{i}'
        for j, ci in ((1, cin), (2, Hh)):
            v = (rng.standard_normal((Hh, ci, 2)) / np.sqrt(2 * ci)).astype(np.float32)
            g = (np.linalg.norm(v.reshape(Hh, -1), axis=1) * (1.0 + 0.2 * rng.standard_normal(Hh))).astype(np.float32)
            sd[f'{q}.conv{j}.bias'] = torch.from_numpy((0.05 * rng.standard_normal(Hh)).astype(np.float32))
            sd[f'{q}.conv{j}.weight_g'] = torch.from_numpy(g.reshape(Hh, 1, 1))
            sd[f'{q}.conv{j}.weight_v'] = torch.from_numpy(v)
        if cin != Hh:
            _lin(sd, rng, q + '.downsample', (Hh, cin, 1), cin)
    _lin(sd, rng, 'text_encoder.decoder', (32, Hh), Hh)
    sd['speaker_embedding.0.weight'] = torch.from_numpy(rng.standard_normal((cfg.n_speakers, cfg.z_size)).astype(np.float32))
    _lin(sd, rng, 'speaker_embedding.1', (cfg.z_size, cfg.z_size), cfg.z_size)
    _lin(sd, rng, 'speaker_mu', (cfg.z_size, cfg.z_size), cfg.z_size)
    _lin(sd, rng, 'speaker_logvar', (cfg.z_size, cfg.z_size), cfg.z_size)
    _gru(sd, rng, 'gru', cfg.gru_in, Hh, cfg.n_layers)
    _lin(sd, rng, 'out.0', (Hh // 2, Hh), Hh)
    _lin(sd, rng, 'out.2', (cfg.pose_dim, Hh // 2), Hh // 2)
    return sd


def discriminator_state_dict(cfg: HotPathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict of ConvDiscriminator (multimodal_context_net.py:207-226)."""
    rng = _rng(2000 + seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    _lin(sd, rng, 'pre_conv.0', (16, cfg.pose_dim, 3), cfg.pose_dim * 3)
    _bn(sd, rng, 'pre_conv.1', 16)
    _lin(sd, rng, 'pre_conv.3', (8, 16, 3), 48)
    _bn(sd, rng, 'pre_conv.4', 8)
    _lin(sd, rng, 'pre_conv.6', (8, 8, 3), 24)
    _gru(sd, rng, 'gru', 8, cfg.d_hidden, cfg.d_layers)
    _lin(sd, rng, 'out', (1, cfg.d_hidden), cfg.d_hidden)
    _lin(sd, rng, 'out2', (1, cfg.n_poses - 6), cfg.n_poses - 6)
    return sd


def embedding_net_state_dict(cfg: HotPathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict of EmbeddingNet(mode='pose') (embedding_net.py:42-82,165-217,262-273)."""
    rng = _rng(3000 + seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    d = cfg.pose_dim
    p = 'pose_encoder'
    for i, (ci, co, k) in enumerate(((d, 32, 3), (32, 64, 3), (64, 64, 4))):
        _lin(sd, rng, f'{p}.net.{i}.0', (co, ci, k), ci * k, scale=1.7)
        _bn(sd, rng, f'{p}.net.{i}.1', co)
    _lin(sd, rng, f'{p}.net.3', (32, 64, 3), 64 * 3, scale=1.7)
    _lin(sd, rng, f'{p}.out_net.0', (256, 384), 384, scale=1.7)
    _bn(sd, rng, f'{p}.out_net.1', 256)
    _lin(sd, rng, f'{p}.out_net.3', (128, 256), 256, scale=1.7)
    _bn(sd, rng, f'{p}.out_net.4', 128)
    _lin(sd, rng, f'{p}.out_net.6', (32, 128), 128, scale=1.7)
    _lin(sd, rng, f'{p}.fc_mu', (32, 32), 32, scale=1.7)
    _lin(sd, rng, f'{p}.fc_logvar', (32, 32), 32)
    q = 'decoder'
    _lin(sd, rng, f'{q}.pre_net.0', (64, 32), 32)
    _bn(sd, rng, f'{q}.pre_net.1', 64)
    _lin(sd, rng, f'{q}.pre_net.3', (136, 64), 64)
    _lin(sd, rng, f'{q}.net.0', (4, 32, 3), 12)          # ConvTranspose1d weight: [Cin, Cout, k]
    sd[f'{q}.net.0.bias'] = torch.from_numpy((0.1 * rng.standard_normal(32)).astype(np.float32))
    _bn(sd, rng, f'{q}.net.1', 32)
    _lin(sd, rng, f'{q}.net.3', (32, 32, 3), 96)
    _bn(sd, rng, f'{q}.net.4', 32)
    _lin(sd, rng, f'{q}.net.6', (32, 32, 3), 96)
    _lin(sd, rng, f'{q}.net.7', (d, 32, 3), 96)
    return sd


def make_inputs(cfg: HotPathConfig, batch: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic TED-shaped clips: audio 0.1*N(0,1) clipped, mostly-PAD word ids
    with 5-9 word-onset frames, random-walk direction vectors, speaker ids U{1..n_spk-1}."""
    rng = _rng(4000 + seed)
    audio = np.clip(0.1 * rng.standard_normal((batch, cfg.audio_len)), -1, 1).astype(np.float32)
    text = np.zeros((batch, cfg.n_poses), dtype=np.int64)
    for b in range(batch):
        n = int(rng.integers(5, 10))
        frames = rng.choice(cfg.n_poses, size=n, replace=False)
        text[b, frames] = rng.integers(4, cfg.n_words, size=n)
    walk = np.cumsum(0.02 * rng.standard_normal((batch, cfg.n_poses, cfg.pose_dim)), axis=1)
    target = (walk + 0.1 * rng.standard_normal((batch, 1, cfg.pose_dim))).astype(np.float32)
    vid = rng.integers(1, cfg.n_speakers, size=batch).astype(np.int64)
    return {'in_audio': torch.from_numpy(audio), 'in_text': torch.from_numpy(text),
            'target': torch.from_numpy(target), 'vid': torch.from_numpy(vid)}


def make_noise(cfg: HotPathConfig, batch: int, seed: int = 0, dropout: bool = False):
    """All random draws of one train_iter_gan call (see trimodal_oracle.StepNoise)."""
    from .trimodal_oracle import StepNoise
    rng = _rng(5000 + seed)
    eps = [torch.from_numpy(rng.standard_normal((batch, cfg.z_size)).astype(np.float32)) for _ in range(3)]
    perm = torch.from_numpy(rng.permutation(batch).astype(np.int64))
    noise = StepNoise(eps=eps, perm=perm)
    if dropout:
        noise.g_masks = [generator_masks(cfg, batch, rng) for _ in range(3)]
        noise.d_masks = [discriminator_masks(cfg, batch, rng) for _ in range(3)]
    return noise


def _mask(rng, shape, p):
    keep = (rng.random(shape) >= p).astype(np.float32) / (1.0 - p)
    return torch.from_numpy(keep)


def generator_masks(cfg: HotPathConfig, batch: int, rng) -> Dict[str, torch.Tensor]:
    T, H = cfg.n_poses, cfg.hidden_size
    m = {'emb': _mask(rng, (batch, T, cfg.wordembed_dim), cfg.emb_dropout)}
    for i in range(cfg.n_layers):
        m[f'tcn{i}_1'] = _mask(rng, (batch, H, T), cfg.dropout_prob)     # oracle layout [B,C,T]
        m[f'tcn{i}_2'] = _mask(rng, (batch, H, T), cfg.dropout_prob)
    for l in range(cfg.n_layers - 1):
        m[f'gru{l}'] = _mask(rng, (batch, T, 2 * H), cfg.dropout_prob)
    return m


def discriminator_masks(cfg: HotPathConfig, batch: int, rng) -> Dict[str, torch.Tensor]:
    T = cfg.n_poses - 6
    return {f'gru{l}': _mask(rng, (batch, T, 2 * cfg.d_hidden), 0.3) for l in range(cfg.d_layers - 1)}


def zeros_like_opt(sd: Dict[str, torch.Tensor]):
    keys = [k for k in sd if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]
    return {'m': {k: torch.zeros_like(sd[k]) for k in keys}, 'v': {k: torch.zeros_like(sd[k]) for k in keys}}


def golden_noise(cfg: HotPathConfig, batch: int, seed: int, use_masks: bool):
    """Noise of the train_e11 / train_e0 golden fixtures: embedding + TCN dropout masks are
    injected, GRU inter-layer dropout is off (the reference GRU cannot take a mask)."""
    noise = make_noise(cfg, batch, seed=seed, dropout=use_masks)
    if use_masks:
        for m in noise.g_masks:
            for l in range(cfg.n_layers - 1):
                m.pop(f'gru{l}', None)
        noise.d_masks = [None, None, None]
    return noise


def with_tcn_aliases(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The reference TemporalBlock registers conv1/conv2 a second time inside ``self.net``
    (tcn.py:30-31), so its state_dict carries ``net.0.*`` / ``net.4.*`` duplicates of
    ``conv1.*`` / ``conv2.*``.  Adds those alias keys (same tensors) for strict loading."""
    out = OrderedDict(sd)
    for k, v in sd.items():
        if '.tcn.network.' in k and '.conv1.' in k:
            out[k.replace('.conv1.', '.net.0.')] = v
        if '.tcn.network.' in k and '.conv2.' in k:
            out[k.replace('.conv2.', '.net.4.')] = v
    return out


# --------------------------------------------------------------------------------------
# seq2seq baseline (config/seq2seq.yml)
# --------------------------------------------------------------------------------------
def _gru_uni(sd, rng, prefix, in_size, hidden, layers):
    s = 1.0 / np.sqrt(hidden)
    for l in range(layers):
        isz = in_size if l == 0 else hidden
        sd[f'{prefix}.weight_ih_l{l}'] = torch.from_numpy((s * rng.standard_normal((3 * hidden, isz))).astype(np.float32))
        sd[f'{prefix}.weight_hh_l{l}'] = torch.from_numpy((s * rng.standard_normal((3 * hidden, hidden))).astype(np.float32))
        sd[f'{prefix}.bias_ih_l{l}'] = torch.from_numpy((s * rng.standard_normal(3 * hidden)).astype(np.float32))
        sd[f'{prefix}.bias_hh_l{l}'] = torch.from_numpy((s * rng.standard_normal(3 * hidden)).astype(np.float32))


def seq2seq_state_dict(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict of Seq2SeqNet (seq2seq_net.py:217-227) in reference key order; cfg = seq2seq_oracle.Seq2SeqConfig."""
    rng = _rng(7000 + seed)
    H, D = cfg.hidden_size, cfg.pose_dim
    sd: Dict[str, torch.Tensor] = OrderedDict()
    sd['encoder.embedding.weight'] = torch.from_numpy((0.5 * rng.standard_normal((cfg.n_words, cfg.wordembed_dim))).astype(np.float32))
    _gru(sd, rng, 'encoder.gru', cfg.wordembed_dim, H, cfg.n_layers)
    p = 'decoder.decoder.'
    _lin(sd, rng, p + 'pre_linear.0', (H, D + H), D + H)
    _bn(sd, rng, p + 'pre_linear.1', H)
    _lin(sd, rng, p + 'attn.attn', (H, 2 * H), 2 * H)
    sd[p + 'attn.v'] = torch.from_numpy((rng.standard_normal(H) / np.sqrt(H)).astype(np.float32))
    _gru_uni(sd, rng, p + 'gru', H, H, cfg.n_layers)
    _lin(sd, rng, p + 'out', (D, H), H)
    return sd


def seq2seq_inputs(cfg, batch: int, seed: int = 0, max_len: int = 12, min_len: int = 4) -> Dict[str, torch.Tensor]:
    """in_text [B, L_max] = [SOS=1, words.., EOS=2] padded with 0 and sorted by decreasing length (what the reference's
    collate function produces, lmdb_data_loader.py:22-41,142-149), lengths [B], target poses [B,n_poses,D]."""
    rng = _rng(8000 + seed)
    lengths = np.sort(rng.integers(min_len, max_len + 1, size=batch))[::-1].copy()
    lengths[0] = max_len
    text = np.zeros((batch, max_len), dtype=np.int64)
    for b in range(batch):
        n = int(lengths[b])
        text[b, 0] = 1
        text[b, 1:n - 1] = rng.integers(4, cfg.n_words, size=n - 2)
        text[b, n - 1] = 2
    walk = np.cumsum(0.02 * rng.standard_normal((batch, cfg.n_poses, cfg.pose_dim)), axis=1)
    target = (walk + 0.1 * rng.standard_normal((batch, 1, cfg.pose_dim))).astype(np.float32)
    return {'in_text': torch.from_numpy(text), 'lengths': torch.from_numpy(lengths.astype(np.int64)), 'target': torch.from_numpy(target)}


def generator_state_dict_variant(cfg, input_context: str, z_mode, seed: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict of PoseGenerator built with args.input_context / z_obj variants (multimodal_context_net.py:72-93): the first GRU
    layer keeps only the input columns that exist ([pre_seq 28 | audio 32 | text 32 | z 16] for 'both' + z), and the speaker layers
    exist only for a Vocab z_obj.  Derived from generator_state_dict so no extra fixture is needed."""
    sd = generator_state_dict(cfg, seed)
    D1 = cfg.pose_dim + 1
    cols = list(range(0, D1))
    if input_context in ('both', 'audio'):
        cols += list(range(D1, D1 + 32))
    if input_context in ('both', 'text'):
        cols += list(range(D1 + 32, D1 + 64))
    if z_mode is not None:
        cols += list(range(D1 + 64, D1 + 64 + cfg.z_size))
    out = OrderedDict()
    for k, v in sd.items():
        if k.startswith('speaker_') and z_mode != 'speaker':
            continue
        if k in ('gru.weight_ih_l0', 'gru.weight_ih_l0_reverse'):
            v = v[:, cols].contiguous()
        out[k] = v
    return out


# --------------------------------------------------------------------------------------
# joint-embedding model: EmbeddingNet(mode != 'pose') (embedding_net.py:130-162,220-273)
# --------------------------------------------------------------------------------------
def joint_embedding_state_dict(cfg: HotPathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict of EmbeddingNet(mode='random'): ContextEncoder + PoseEncoderConv + PoseDecoderGRU, reference key order."""
    rng = _rng(9000 + seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    g = generator_state_dict(cfg, seed + 5)
    for k, v in g.items():                                    # ContextEncoder.text_encoder / .audio_encoder (:225-226)
        if k.startswith('text_encoder.'):
            sd['context_encoder.' + k] = v
    for k, v in g.items():
        if k.startswith('audio_encoder.'):
            sd['context_encoder.' + k] = v
    _gru_uni(sd, rng, 'context_encoder.gru', 64, 256, 2)       # :227-228
    _lin(sd, rng, 'context_encoder.out.0', (128, 256), 256, scale=1.5)
    _bn(sd, rng, 'context_encoder.out.1', 128)
    _lin(sd, rng, 'context_encoder.out.3', (32, 128), 128, scale=1.5)
    _lin(sd, rng, 'context_encoder.fc_mu', (32, 32), 32, scale=1.5)
    _lin(sd, rng, 'context_encoder.fc_logvar', (32, 32), 32, scale=0.5)
    pe = embedding_net_state_dict(cfg, seed + 3)
    for k, v in pe.items():
        if k.startswith('pose_encoder.'):
            sd[k] = v
    _lin(sd, rng, 'decoder.pre_pose_net.0', (32, cfg.pose_dim * 4), cfg.pose_dim * 4, scale=1.5)      # PoseDecoderGRU :138-143
    _bn(sd, rng, 'decoder.pre_pose_net.1', 32)
    _lin(sd, rng, 'decoder.pre_pose_net.3', (32, 32), 32, scale=1.5)
    _gru(sd, rng, 'decoder.gru', 64, 300, 4)
    _lin(sd, rng, 'decoder.out.0', (150, 300), 300)
    _lin(sd, rng, 'decoder.out.2', (cfg.pose_dim, 150), 150)
    return sd



def s2g_state_dict(template, seed):
    """Deterministic Speech2Gesture weights (tests / goldens): a value for every entry of `template` (a state_dict of the reference's or
    of our Generator / Discriminator - same keys and shapes), drawn from a seeded generator in key order, fan-in scaled."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in template.items():
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(v)
        elif k.endswith('running_mean'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('running_var'):
            out[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif v.dim() == 1 and k.endswith('weight'):                 # BatchNorm gamma
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:                                          # biases, BatchNorm beta
            out[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (1.4 / fan_in ** 0.5)
    return out
