"""Drop-in PoseGenerator / ConvDiscriminator / WavEncoder / TextEncoderTCN for the B200.

Same constructors, forward signatures, public attributes and state_dict keys as the reference
(scripts/model/multimodal_context_net.py:9-28, 31-61, 64-160, 207-252) so that train.py:42-48,188-190,281,389,
synthesize.py:131 and train_utils.py:152-183 keep working and reference checkpoints load with strict=True.
The standard torch sub-modules below are PARAMETER CONTAINERS ONLY (identical default initialisation, identical key
names); none of their forward()s is ever called.  All arithmetic runs in hand-written sm_100a kernels through the C ABI
(tgb200.engine); CPU tensors raise - there is no fallback path."""
import itertools

import torch
import torch.nn as nn

from model import vocab
from model.tcn import TemporalConvNet
from tgb200 import _lib, config, ops
from tgb200.engine import DiscriminatorEngine, GeneratorEngine

_RING = 4          # live training forwards whose activations are kept for a later backward (reference pattern needs 3)


class WavEncoder(nn.Module):
    """Raw 16 kHz audio -> 32-d feature per frame (multimodal_context_net.py:9-28)."""

    def __init__(self):
        super().__init__()
        self.feat_extractor = nn.Sequential(
            nn.Conv1d(1, 16, 15, stride=5, padding=1600), nn.BatchNorm1d(16), nn.LeakyReLU(0.3, inplace=True),
            nn.Conv1d(16, 32, 15, stride=6), nn.BatchNorm1d(32), nn.LeakyReLU(0.3, inplace=True),
            nn.Conv1d(32, 64, 15, stride=6), nn.BatchNorm1d(64), nn.LeakyReLU(0.3, inplace=True),
            nn.Conv1d(64, 32, 15, stride=6))

    def forward(self, wav_data):
        """[B, L] raw audio -> [B, n_frames, 32] (multimodal_context_net.py:25-28).  Stand-alone call: forward only, no autograd graph; inside
        PoseGenerator / ContextEncoder the parent's engine runs the same kernels together with their backward."""
        _lib.require_cuda()
        if not wav_data.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('WavEncoder.forward runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        if getattr(self, '_engine', None) is None:
            from tgb200.engine import StandaloneEncoderEngine
            self._engine = StandaloneEncoderEngine(self, 'audio_encoder.')
        return self._engine.ensure(wav_data.device).run_wav(wav_data)


class TextEncoderTCN(nn.Module):
    """Word ids -> 32-d feature per frame (multimodal_context_net.py:31-61)."""

    def __init__(self, args, n_words, embed_size=300, pre_trained_embedding=None, kernel_size=2, dropout=0.3, emb_dropout=0.1):
        super().__init__()
        if pre_trained_embedding is not None:
            assert pre_trained_embedding.shape[0] == n_words
            assert pre_trained_embedding.shape[1] == embed_size
            self.embedding = nn.Embedding.from_pretrained(torch.FloatTensor(pre_trained_embedding), freeze=args.freeze_wordembed)
        else:
            self.embedding = nn.Embedding(n_words, embed_size)
        num_channels = [args.hidden_size] * args.n_layers
        self.tcn = TemporalConvNet(embed_size, num_channels, kernel_size, dropout=dropout)
        self.decoder = nn.Linear(num_channels[-1], 32)
        self.drop = nn.Dropout(emb_dropout)
        self.emb_dropout = emb_dropout
        self.init_weights()

    def init_weights(self):
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.normal_(0, 0.01)

    def forward(self, input):
        """[B, T] word ids -> ([B, T, 32], 0) (multimodal_context_net.py:57-61; train mode draws its dropout masks from Philox).  Stand-alone
        call: forward only, no autograd graph; inside PoseGenerator / ContextEncoder the parent's engine runs the same kernels + backward."""
        _lib.require_cuda()
        if not input.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('TextEncoderTCN.forward runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        if getattr(self, '_engine', None) is None:
            from tgb200.engine import StandaloneEncoderEngine
            self._engine = StandaloneEncoderEngine(self, 'text_encoder.')
            self._noise = _NoiseSource(_module_seed() ^ 0x7E47)
        eng = self._engine.ensure(input.device)
        y = eng.run_text(input, self._noise.seed, self._noise.offset_dev(input.device))
        if self.training:
            self._noise.advance()
        return y, 0


class _NoiseSource:
    """Philox stream state of one module: a device-resident offset so CUDA-graph replays draw fresh numbers."""

    def __init__(self, seed):
        self.base_seed = seed
        self.offset = None

    @property
    def seed(self):
        """Data-parallel ranks draw INDEPENDENT dropout masks / reparameterisation noise / speaker permutations (like the per-replica
        generators of nn.DataParallel, train.py:93-96): the rank is mixed into the Philox key at use time, so identical weights no
        longer have to come from identical RNG seeds (the flat arena is broadcast from rank 0 once, train_gan._dp_sync_once)."""
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        return (self.base_seed ^ (rank * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF

    def offset_dev(self, device):
        if self.offset is None or self.offset.device != device:
            self.offset = torch.zeros(1, dtype=torch.int64, device=device)
        return self.offset

    def advance(self):
        if self.offset is not None:
            ops.increment_i64(self.offset, 1)


def _module_seed():
    return int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF) ^ 0x5DEECE66D


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, pre_seq, in_text, in_audio, vid, eps, masks, slot, *params):
        eng = module.engine().ensure(pre_seq.device, slot)
        eng.prep_weights()
        B = pre_seq.shape[0]
        training = module.training
        poses, z, mu, logvar = eng.forward(pre_seq, in_text, in_audio, vid, eps, B, training, masks,
                                           n_bn_updates=1, save=torch.is_grad_enabled() or True)
        ctx.module, ctx.slot, ctx.B, ctx.device = module, slot, B, pre_seq.device
        ctx.gen = module._slot_gen[slot] = next(module._gen_counter)
        ctx.fwd_ctx = eng.ctx
        outs = [poses.clone()]
        for t in (z, mu, logvar):
            outs.append(t.clone() if t is not None else None)
        ctx.has = [t is not None for t in (z, mu, logvar)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_poses, d_z, d_mu, d_logvar):
        module = ctx.module
        if module._slot_gen.get(ctx.slot) != ctx.gen:
            raise RuntimeError('PoseGenerator activations of this forward were recycled (more than %d training forwards '
                               'were kept alive before backward)' % _RING)
        eng = module.engine().ensure(ctx.device, ctx.slot)
        eng.ctx = ctx.fwd_ctx
        B = ctx.B
        c = lambda t: t.contiguous().clone() if t is not None else None
        eng.backward(d_poses.contiguous(), 0, B, d_mu=c(d_mu), d_logvar=c(d_logvar), d_z=(d_z.contiguous() if d_z is not None else None))
        return (None,) * (8 + len(module._param_list))


class PoseGenerator(nn.Module):
    def __init__(self, args, pose_dim, n_words, word_embed_size, word_embeddings, z_obj=None):
        super().__init__()
        self.pre_length = args.n_pre_poses
        self.gen_length = args.n_poses - args.n_pre_poses
        self.z_obj = z_obj
        self.input_context = args.input_context
        self.pose_dim = pose_dim
        if self.input_context == 'both':
            self.in_size = 32 + 32 + pose_dim + 1      # audio_feat + text_feat + last pose + constraint bit
        elif self.input_context == 'none':
            self.in_size = pose_dim + 1
        else:
            self.in_size = 32 + pose_dim + 1
        self.audio_encoder = WavEncoder()
        self.text_encoder = TextEncoderTCN(args, n_words, word_embed_size, pre_trained_embedding=word_embeddings, dropout=args.dropout_prob)
        self.speaker_embedding = None
        self.z_mode = None
        if self.z_obj:
            self.z_size = 16
            self.in_size += self.z_size
            if isinstance(self.z_obj, vocab.Vocab) or type(self.z_obj).__name__ == 'Vocab':
                self.speaker_embedding = nn.Sequential(nn.Embedding(z_obj.n_words, self.z_size), nn.Linear(self.z_size, self.z_size))
                self.speaker_mu = nn.Linear(self.z_size, self.z_size)
                self.speaker_logvar = nn.Linear(self.z_size, self.z_size)
                self.z_mode = 'speaker'
            else:
                self.z_mode = 'random'
        self.hidden_size = args.hidden_size
        self.gru = nn.GRU(self.in_size, hidden_size=self.hidden_size, num_layers=args.n_layers, batch_first=True, bidirectional=True,
                          dropout=args.dropout_prob)
        self.out = nn.Sequential(nn.Linear(self.hidden_size, self.hidden_size // 2), nn.LeakyReLU(True),
                                 nn.Linear(self.hidden_size // 2, pose_dim))
        self.do_flatten_parameters = False          # kept for attribute compatibility; there is no cuDNN weight buffer here
        self._engine = None
        self._noise = _NoiseSource(_module_seed())
        self._injected = None
        self._slot_gen = {}
        self._gen_counter = itertools.count(1)
        self._ring = itertools.cycle(range(_RING))
        self._graphs = {}

    # ---- engine plumbing (not part of the reference API) ------------------------------------------------------------
    def engine(self) -> GeneratorEngine:
        if self._engine is None:
            self._engine = GeneratorEngine(self)
            self._param_list = [p for p in self.parameters()]
        return self._engine

    def set_noise(self, eps=None, masks=None):
        """Test seam: the next forward uses this reparameterisation noise eps [B,16] and these dropout keep-masks
        (channels-last, already scaled by 1/(1-p)) instead of drawing them with the Philox kernels."""
        self._injected = (eps, masks)

    def forward(self, pre_seq, in_text, in_audio, vid_indices=None):
        _lib.require_cuda()
        if not pre_seq.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('PoseGenerator runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        dev = pre_seq.device
        eng = self.engine()
        B, T = pre_seq.shape[0], pre_seq.shape[1]
        if self.input_context != 'none':
            assert in_audio is not None and in_text is not None
        if self.z_mode == 'speaker':
            assert vid_indices is not None
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        slot = ('ring%d_%d' % (next(self._ring), B)) if grad else ('nograd_%d' % B)
        if not grad and self._injected is None and config.graphs() and not _lib.TRACE_ONLY:
            out = self._forward_graphed(pre_seq, in_text, in_audio, vid_indices, slot)
            if out is not None:
                return out
        eng.ensure(dev, slot)
        eps, masks = (None, None)
        if self._injected is not None:
            eps, masks = self._injected
            self._injected = None
        off = self._noise.offset_dev(dev)
        if eps is None and self.z_mode is not None:
            eps = eng.ws.get('noise.eps', (B, 16))
            ops.philox_normal(eps, B * 16, self._noise.seed, off, 1000)
        if masks is None and self.training:
            masks = eng.make_masks(B, T, self._noise.seed, off)
        if not self.training:
            masks = None
        self._noise.advance()
        pre_seq = pre_seq.contiguous().float()
        in_text_c = in_text.contiguous() if in_text is not None else None
        in_audio_c = in_audio.contiguous().float() if in_audio is not None else None
        vid = vid_indices.contiguous() if vid_indices is not None else None
        if eps is not None:
            eps = eps.contiguous()
        outs = _GeneratorFn.apply(self, pre_seq, in_text_c, in_audio_c, vid, eps, masks, slot, *self._param_list)
        poses, z, mu, logvar = outs
        if in_text_c is not None and self.input_context != 'none':
            assert poses.shape[1] == in_text.shape[1]
        return poses, z, mu, logvar


def _pg_forward_graphed(self, pre_seq, in_text, in_audio, vid_indices, slot):
    """Inference fast path (no autograd): after two eager calls with a given shape the ~70 launches of the forward are
    captured into a CUDA graph on static input buffers and replayed; the speaker-style noise is drawn inside the graph
    from the device-resident Philox offset, so every call still samples a fresh z (embedding_net.py:10-13)."""
    dev = pre_seq.device
    eng = self.engine().ensure(dev, slot)
    B, T = pre_seq.shape[0], pre_seq.shape[1]
    key = (slot, dev.index, tuple(pre_seq.shape), tuple(in_text.shape) if in_text is not None else None,
           tuple(in_audio.shape) if in_audio is not None else None, self.training, config.mode(), config.overlap())
    st = self._graphs.setdefault(key, {'calls': 0, 'graph': None, 'failed': False})
    st['calls'] += 1
    if st['failed'] or st['calls'] <= 2:
        return None
    ws = eng.ws
    if 'pre' not in st:
        st['pre'] = ws.get('gi.pre', tuple(pre_seq.shape))
        st['text'] = ws.get('gi.text', tuple(in_text.shape), torch.int64) if in_text is not None else None
        st['audio'] = ws.get('gi.audio', tuple(in_audio.shape)) if in_audio is not None else None
        st['vid'] = ws.get('gi.vid', tuple(vid_indices.shape), torch.int64) if vid_indices is not None else None
    st['pre'].copy_(pre_seq)
    if st['text'] is not None:
        st['text'].copy_(in_text)
    if st['audio'] is not None:
        st['audio'].copy_(in_audio)
    if st['vid'] is not None:
        st['vid'].copy_(vid_indices)
    if st['graph'] is None:
        try:
            off = self._noise.offset_dev(dev)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                eps = None
                if self.z_mode is not None:
                    eps = ws.get('noise.eps', (B, 16))
                    ops.philox_normal(eps, B * 16, self._noise.seed, off, 1000)
                masks = eng.make_masks(B, T, self._noise.seed, off) if self.training else None
                self._noise.advance()
                eng.prep_weights()
                st['outs'] = eng.forward(st['pre'], st['text'], st['audio'], st['vid'], eps, B, self.training, masks, n_bn_updates=1,
                                         save=False)
            st['graph'] = graph
        except Exception as exc:
            st['failed'] = True
            import warnings
            warnings.warn('tgb200: CUDA-graph capture of PoseGenerator.forward failed (%s); using eager launches' % (str(exc).splitlines()[0],))
            torch.cuda.synchronize()
            return None
    st['graph'].replay()
    poses, z, mu, logvar = st['outs']
    return (poses.clone(), z.clone() if z is not None else None, mu.clone() if mu is not None else None,
            logvar.clone() if logvar is not None else None)


PoseGenerator._forward_graphed = _pg_forward_graphed


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, poses, masks, slot, *params):
        eng = module.engine().ensure(poses.device, slot)
        eng.prep_weights()
        prob = eng.forward(poses, module.training, masks, save=True).clone()
        ctx.module, ctx.slot, ctx.device = module, slot, poses.device
        ctx.gen = module._slot_gen[slot] = next(module._gen_counter)
        ctx.fwd_ctx = eng.ctx
        ctx.fwd_Ts = list(eng.Ts)
        ctx.need_dposes = poses.requires_grad
        ctx.save_for_backward(prob)
        return prob

    @staticmethod
    def backward(ctx, d_prob):
        module = ctx.module
        if module._slot_gen.get(ctx.slot) != ctx.gen:
            raise RuntimeError('ConvDiscriminator activations of this forward were recycled before backward')
        (prob,) = ctx.saved_tensors
        eng = module.engine().ensure(ctx.device, ctx.slot)
        eng.ctx, eng.Ts = ctx.fwd_ctx, ctx.fwd_Ts
        dlogit = (d_prob * prob * (1.0 - prob)).contiguous()
        dposes = eng.backward(dlogit, ctx.need_dposes)
        return (None, dposes.clone() if dposes is not None else None, None, None) + (None,) * len(module._param_list)


class ConvDiscriminator(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = 64
        self.pre_conv = nn.Sequential(
            nn.Conv1d(input_size, 16, 3), nn.BatchNorm1d(16), nn.LeakyReLU(True),
            nn.Conv1d(16, 8, 3), nn.BatchNorm1d(8), nn.LeakyReLU(True),
            nn.Conv1d(8, 8, 3))
        self.gru = nn.GRU(8, hidden_size=self.hidden_size, num_layers=4, bidirectional=True, dropout=0.3, batch_first=True)
        self.out = nn.Linear(self.hidden_size, 1)
        self.out2 = nn.Linear(28, 1)
        self.do_flatten_parameters = False
        self._engine = None
        self._noise = _NoiseSource(_module_seed() ^ 0xD15C)
        self._injected = None
        self._slot_gen = {}
        self._gen_counter = itertools.count(1)
        self._ring = itertools.cycle(range(_RING))

    def engine(self) -> DiscriminatorEngine:
        if self._engine is None:
            self._engine = DiscriminatorEngine(self)
            self._param_list = [p for p in self.parameters()]
        return self._engine

    def set_noise(self, masks=None):
        self._injected = masks

    def forward(self, poses, in_text=None):
        _lib.require_cuda()
        if not poses.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('ConvDiscriminator runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        eng = self.engine()
        B = poses.shape[0]
        grad = torch.is_grad_enabled()
        slot = ('ring%d_%d' % (next(self._ring), B)) if grad else ('nograd_%d' % B)
        eng.ensure(poses.device, slot)
        masks = self._injected
        self._injected = None
        if masks is None and self.training:
            masks = eng.make_masks(B, poses.shape[1] - 6, self._noise.seed, self._noise.offset_dev(poses.device))
            self._noise.advance()
        if not self.training:
            masks = None
        return _DiscriminatorFn.apply(self, poses.contiguous().float(), masks, slot, *self._param_list)
