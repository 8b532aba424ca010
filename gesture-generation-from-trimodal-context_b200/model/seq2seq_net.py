"""Drop-in Seq2SeqNet (text -> gesture baseline) for the B200: same classes, constructors, forward signatures and
state_dict keys as the reference (scripts/model/seq2seq_net.py:14-254), so train.py:55-57,197-198 and reference checkpoints
(strict=True) keep working.  The torch sub-modules are PARAMETER CONTAINERS ONLY; the arithmetic runs in hand-written
sm_100a kernels through the C ABI (tgb200.seq2seq_engine).  CPU tensors raise - there is no fallback path.

Training goes through train_eval.train_seq2seq.train_iter_seq2seq (the reference's only training call site); calling
forward() directly returns the generated poses without an autograd graph."""
import math

import torch
import torch.nn as nn

from tgb200 import _lib
from tgb200.seq2seq_engine import Seq2SeqEngine


class EncoderRNN(nn.Module):
    """seq2seq_net.py:14-59."""

    def __init__(self, input_size, embed_size, hidden_size, n_layers=1, dropout=0.5, pre_trained_embedding=None):
        super().__init__()
        self.input_size, self.hidden_size, self.embed_size, self.n_layers, self.dropout = input_size, hidden_size, embed_size, n_layers, dropout
        if pre_trained_embedding is not None:
            assert pre_trained_embedding.shape[0] == input_size
            assert pre_trained_embedding.shape[1] == embed_size
            self.embedding = nn.Embedding.from_pretrained(torch.FloatTensor(pre_trained_embedding), freeze=False)
        else:
            self.embedding = nn.Embedding(input_size, embed_size)
        self.gru = nn.GRU(embed_size, hidden_size, n_layers, dropout=self.dropout, bidirectional=True)
        self.do_flatten_parameters = False

    def forward(self, input_seqs, input_lengths, hidden=None):
        raise RuntimeError('EncoderRNN holds parameters only; its kernels run inside Seq2SeqNet.forward (no PyTorch fallback)')


class Attn(nn.Module):
    """seq2seq_net.py:62-94."""

    def __init__(self, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.attn = nn.Linear(self.hidden_size * 2, hidden_size)
        self.v = nn.Parameter(torch.rand(hidden_size))
        stdv = 1. / math.sqrt(self.v.size(0))
        self.v.data.normal_(mean=0, std=stdv)

    def forward(self, hidden, encoder_outputs):
        raise RuntimeError('Attn holds parameters only; its kernels run inside Seq2SeqNet.forward (no PyTorch fallback)')


class BahdanauAttnDecoderRNN(nn.Module):
    """seq2seq_net.py:97-198."""

    def __init__(self, input_size, hidden_size, output_size, n_layers=1, dropout_p=0.1, discrete_representation=False, speaker_model=None):
        super().__init__()
        self.hidden_size, self.output_size, self.n_layers, self.dropout_p = hidden_size, output_size, n_layers, dropout_p
        self.discrete_representation, self.speaker_model = discrete_representation, speaker_model
        if self.discrete_representation:
            self.embedding = nn.Embedding(output_size, hidden_size)
            self.dropout = nn.Dropout(dropout_p)
        if self.speaker_model:
            self.speaker_embedding = nn.Embedding(speaker_model.n_words, 8)
        if self.discrete_representation:
            input_size = hidden_size
        linear_input_size = input_size + hidden_size
        if self.speaker_model:
            linear_input_size += 8
        self.pre_linear = nn.Sequential(nn.Linear(linear_input_size, hidden_size), nn.BatchNorm1d(hidden_size), nn.ReLU(inplace=True))
        self.attn = Attn(hidden_size)
        self.gru = nn.GRU(hidden_size, hidden_size, n_layers, dropout=dropout_p)
        self.out = nn.Linear(hidden_size, output_size)
        self.do_flatten_parameters = False

    def freeze_attn(self):
        for param in self.attn.parameters():
            param.requires_grad = False

    def forward(self, motion_input, last_hidden, encoder_outputs, vid_indices=None):
        raise RuntimeError('BahdanauAttnDecoderRNN holds parameters only; its kernels run inside Seq2SeqNet.forward (no PyTorch fallback)')


class Generator(nn.Module):
    """seq2seq_net.py:201-214."""

    def __init__(self, args, motion_dim, discrete_representation=False, speaker_model=None):
        super().__init__()
        self.output_size = motion_dim
        self.n_layers = args.n_layers
        self.discrete_representation = discrete_representation
        self.decoder = BahdanauAttnDecoderRNN(input_size=motion_dim + args.GAN_noise_size, hidden_size=args.hidden_size,
                                              output_size=self.output_size, n_layers=self.n_layers, dropout_p=args.dropout_prob,
                                              discrete_representation=discrete_representation, speaker_model=speaker_model)

    def freeze_attn(self):
        self.decoder.freeze_attn()

    def forward(self, z, motion_input, last_hidden, encoder_output, vid_indices=None):
        raise RuntimeError('Generator holds parameters only; its kernels run inside Seq2SeqNet.forward (no PyTorch fallback)')


class _Seq2SeqFn(torch.autograd.Function):
    """Train-mode Seq2SeqNet under torch autograd (the reference's own loop, train_seq2seq.py:39-51: custom_loss in torch, loss.backward(),
    clip_grad_norm_, optim.step()): forward and hand-derived backward are the launch plans of Seq2SeqEngine; parameter gradients are
    accumulated into the .grad views of the flat arena."""

    @staticmethod
    def forward(ctx, module, in_text, lens_dev, Tm, poses, masks, slot, *params):
        eng = module.engine().ensure(poses.device, slot)
        out = eng.forward(in_text, lens_dev, Tm, poses, True, masks, save=True)
        ctx.module, ctx.device, ctx.slot, ctx.fwd_ctx = module, poses.device, slot, eng.ctx
        ctx.gen = module._fwd_gen = getattr(module, '_fwd_gen', 0) + 1
        return out.clone()

    @staticmethod
    def backward(ctx, d_out):
        module = ctx.module
        if module._fwd_gen != ctx.gen:
            raise RuntimeError('Seq2SeqNet activations of this forward were overwritten by a later training forward before backward')
        eng = module.engine().ensure(ctx.device, ctx.slot)
        eng.ctx = ctx.fwd_ctx
        eng.backward_from_dy(d_out.transpose(0, 1).contiguous())          # the backward sweep walks time: [T,B,D]
        return (None,) * (7 + len(module._fn_params))


class Seq2SeqNet(nn.Module):
    """seq2seq_net.py:217-254."""

    def __init__(self, args, pose_dim, n_frames, n_words, word_embed_size, word_embeddings, speaker_model=None):
        super().__init__()
        assert getattr(args, 'GAN_noise_size', 0) == 0, 'GAN_noise_size > 0 is not on the configured path (parse_args.py:53 default 0)'
        self.encoder = EncoderRNN(n_words, word_embed_size, args.hidden_size, args.n_layers, dropout=args.dropout_prob,
                                  pre_trained_embedding=word_embeddings)
        self.decoder = Generator(args, pose_dim, speaker_model=speaker_model)
        self.n_frames = n_frames
        self.n_pre_poses = args.n_pre_poses
        self._engine = None
        from model.multimodal_context_net import _NoiseSource, _module_seed
        self._noise = _NoiseSource(_module_seed())

    def engine(self) -> Seq2SeqEngine:
        if self._engine is None:
            self._engine = Seq2SeqEngine(self)
        return self._engine

    @staticmethod
    def prepare_lengths(in_lengths, device):
        """in_lengths: list / CPU tensor (what pack_padded_sequence requires, seq2seq_net.py:52) -> (device int64 [B], max length)."""
        lens = torch.as_tensor(in_lengths, dtype=torch.int64).cpu()
        assert bool((lens[:-1] >= lens[1:]).all()), 'in_lengths must be sorted in decreasing order (pack_padded_sequence, seq2seq_net.py:52)'
        return lens.to(device), int(lens.max())

    def forward(self, in_text, in_lengths, poses, vid_indices):
        _lib.require_cuda()
        if not poses.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('Seq2SeqNet runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        dev = poses.device
        lens_dev, Tm = self.prepare_lengths(in_lengths, dev)
        grad = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        slot = ('autograd_%d_%d' % (poses.shape[0], Tm)) if grad else ('fwd_%d' % poses.shape[0])
        eng = self.engine().ensure(dev, slot)
        masks = None
        if self.training:
            masks = eng.make_masks(poses.shape[0], Tm, self._noise.seed, self._noise.offset_dev(dev))
            self._noise.advance()
        if grad:
            self._fn_params = [p for p in self.parameters() if p.requires_grad]
            return _Seq2SeqFn.apply(self, in_text.contiguous(), lens_dev, Tm, poses.detach().contiguous().float(), masks, slot, *self._fn_params)
        out = eng.forward(in_text.contiguous(), lens_dev, Tm, poses.contiguous().float(), self.training, masks, save=False)
        return out.clone()
