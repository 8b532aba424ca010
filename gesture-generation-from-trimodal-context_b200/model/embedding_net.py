"""EmbeddingNet(mode='pose') - the pose auto-encoder whose 32-d latent is the feature space of the Frechet Gesture
Distance (reference: scripts/model/embedding_net.py:10-13, 16-39, 42-82, 165-217, 262-314).

Parameter containers with the reference's names.  Eval mode (what FGD uses): 14 fused sm_100a GEMM launches with the
running-statistics BatchNorm folded into their epilogues (tgb200.engine.EmbeddingEngine).  Train mode (batch statistics,
running-statistics update): tgb200.embed_engine.AutoEncoderTrainEngine, which also holds the hand-derived backward used by the
auto-encoder step functions (train_feature_extractor.train_iter, train_eval.train_joint_embed.train_iter_embed).
The module-level forward returns detached tensors: training goes through those step functions, not through autograd.
mode != 'pose' builds the joint-embedding model (ContextEncoder + PoseEncoderConv + PoseDecoderGRU, embedding_net.py:130-162,220-273;
launch plan tgb200.embed_engine.JointEmbeddingEngine, step function train_eval.train_joint_embed.train_iter_embed)."""
import random

import torch
import torch.nn as nn

from tgb200 import _lib, ops
from tgb200.embed_engine import AutoEncoderTrainEngine, JointEmbeddingEngine
from tgb200.engine import EmbeddingEngine


class _AutoEncoderFn(torch.autograd.Function):
    """Train-mode EmbeddingNet(mode='pose') under torch autograd (the reference's own loop: loss.backward(); optim.step(),
    train_feature_extractor.py:54-97): forward and hand-derived backward are the launch plans of AutoEncoderTrainEngine; parameter
    gradients are accumulated straight into the .grad views of the flat arena, like torch accumulates into .grad."""

    @staticmethod
    def forward(ctx, module, poses, *params):
        eng = module.train_engine().ensure(poses.device)
        mu, logvar, recon = eng.forward(poses, training=True)
        ctx.module, ctx.device, ctx.fwd_ctx, ctx.shapes = module, poses.device, eng.ctx, (eng.enc_T, eng.dec_T)
        ctx.gen = module._ae_gen = getattr(module, '_ae_gen', 0) + 1
        return mu.clone(), mu.clone(), logvar.clone(), recon.clone()

    @staticmethod
    def backward(ctx, d_feat, d_mu, d_logvar, d_recon):
        module = ctx.module
        if module._ae_gen != ctx.gen:
            raise RuntimeError('EmbeddingNet activations of this forward were overwritten by a later training forward before backward')
        eng = module.train_engine().ensure(ctx.device)
        eng.ctx, (eng.enc_T, eng.dec_T) = ctx.fwd_ctx, ctx.shapes
        extra = None
        for g in (d_feat, d_mu):                       # poses_feat and pose_mu are the same tensor when variational_encoding is False
            if g is not None:
                extra = g.contiguous().clone() if extra is None else extra + g
        # pose_logvar: fc_logvar's backward is not part of this plan (the reference never back-propagates through it, :58-86)
        assert d_logvar is None or not bool(d_logvar.any()), 'a loss on pose_logvar is not supported by the auto-encoder plan'
        B, T, D = eng.ctx['B'], eng.ctx['T'], eng.ctx['D']
        eng.backward(d_recon.contiguous().view(B * T, D) if d_recon is not None else None, extra)
        return (None,) * (2 + len(module._ae_params))


class _JointEmbeddingFn(torch.autograd.Function):
    """Train-mode joint-embedding EmbeddingNet under torch autograd (the reference's loop, train_joint_embed.py:5-51).  Only the
    reconstruction carries a gradient there; the decoder and the encoder whose latent was decoded receive it.  The other encoder's
    parameters get NO gradient in the reference (.grad stays None and torch.optim.Adam skips them): their .grad views are dropped again
    after the backward so that the caller's optimiser behaves the same."""

    @staticmethod
    def forward(ctx, module, in_text, in_audio, pre_poses, poses, branch, *params):
        eng = module.joint_engine().ensure(pre_poses.device)
        r = eng.forward(in_text, in_audio, pre_poses, poses, branch, True)
        ctx.module, ctx.device, ctx.st, ctx.pose_ctx = module, pre_poses.device, eng.st, getattr(eng.pose, 'ctx', None)
        ctx.gen = module._joint_gen = getattr(module, '_joint_gen', 0) + 1
        ctx.present = [r[k] is not None for k in ('c_feat', 'c_mu', 'c_lv', 'p_mu', 'p_mu', 'p_lv')]
        outs = [r[k].clone() if r[k] is not None else None for k in ('c_feat', 'c_mu', 'c_lv', 'p_mu', 'p_mu', 'p_lv')]
        return tuple(outs) + (r['out'].clone(),)

    @staticmethod
    def backward(ctx, d_cf, d_cmu, d_clv, d_pf, d_pmu, d_plv, d_out):
        module = ctx.module
        if module._joint_gen != ctx.gen:
            raise RuntimeError('EmbeddingNet activations of this forward were overwritten by a later training forward before backward')
        for g in (d_cf, d_cmu, d_clv, d_pf, d_pmu, d_plv):
            assert g is None or not bool(g.any()), 'only a loss on the reconstruction is supported by the joint-embedding plan'
        eng = module.joint_engine().ensure(ctx.device)
        eng.st = ctx.st
        if ctx.pose_ctx is not None:
            eng.pose.ctx = ctx.pose_ctx
        B = eng.st['B']
        eng.backward(d_out.contiguous().view(B * eng.T, eng.D))
        unused = module.pose_encoder if eng.st['branch'] == 'speech' else module.context_encoder
        for p in unused.parameters():
            p.grad = None
        return (None,) * (6 + len(module._joint_params))


def reparameterize(mu, logvar):
    """embedding_net.py:10-13: mu + eps*exp(0.5*logvar), eps from the Philox kernel (CUDA tensors only)."""
    _lib.require_cuda()
    mu_c, lv_c = mu.detach().contiguous().float(), logvar.detach().contiguous().float()
    eps = torch.empty_like(mu_c)
    off = torch.zeros(1, dtype=torch.int64, device=mu.device)
    ops.philox_normal(eps, eps.numel(), int(torch.randint(0, 2 ** 62, (1,)).item()), off, 7)
    z = torch.empty_like(mu_c)
    ops.reparam_fwd(mu_c, lv_c, eps, z, z.numel())
    return z


def ConvNormRelu(in_channels, out_channels, downsample=False, padding=0, batchnorm=True):
    k, s = (4, 2) if downsample else (3, 1)
    conv_block = nn.Conv1d(in_channels, out_channels, kernel_size=k, stride=s, padding=padding)
    if batchnorm:
        return nn.Sequential(conv_block, nn.BatchNorm1d(out_channels), nn.LeakyReLU(0.2, True))
    return nn.Sequential(conv_block, nn.LeakyReLU(0.2, True))


class PoseEncoderConv(nn.Module):
    def __init__(self, length, dim):
        super().__init__()
        self.net = nn.Sequential(ConvNormRelu(dim, 32, batchnorm=True), ConvNormRelu(32, 64, batchnorm=True),
                                 ConvNormRelu(64, 64, True, batchnorm=True), nn.Conv1d(64, 32, 3))
        self.out_net = nn.Sequential(nn.Linear(384, 256), nn.BatchNorm1d(256), nn.LeakyReLU(True),      # 384: 34-frame clips
                                     nn.Linear(256, 128), nn.BatchNorm1d(128), nn.LeakyReLU(True), nn.Linear(128, 32))
        self.fc_mu = nn.Linear(32, 32)
        self.fc_logvar = nn.Linear(32, 32)

    def forward(self, poses, variational_encoding):
        raise RuntimeError('PoseEncoderConv holds parameters only; run EmbeddingNet.forward')


class PoseDecoderConv(nn.Module):
    def __init__(self, length, dim, use_pre_poses=False):
        super().__init__()
        assert not use_pre_poses, 'use_pre_poses=True is never constructed by the reference (embedding_net.py:273)'
        self.use_pre_poses = use_pre_poses
        feat_size = 32
        if length == 64:
            self.pre_net = nn.Sequential(nn.Linear(feat_size, 128), nn.BatchNorm1d(128), nn.LeakyReLU(True), nn.Linear(128, 256))
        elif length == 34:
            self.pre_net = nn.Sequential(nn.Linear(feat_size, 64), nn.BatchNorm1d(64), nn.LeakyReLU(True), nn.Linear(64, 136))
        else:
            assert False
        self.net = nn.Sequential(nn.ConvTranspose1d(4, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
                                 nn.ConvTranspose1d(32, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
                                 nn.Conv1d(32, 32, 3), nn.Conv1d(32, dim, 3))

    def forward(self, feat, pre_poses=None):
        raise RuntimeError('PoseDecoderConv holds parameters only; run EmbeddingNet.forward')


class PoseDecoderGRU(nn.Module):
    """embedding_net.py:130-162 (parameter container)."""

    def __init__(self, gen_length, pose_dim):
        super().__init__()
        self.gen_length = gen_length
        self.pose_dim = pose_dim
        self.in_size = 32 + 32
        self.hidden_size = 300
        self.pre_pose_net = nn.Sequential(nn.Linear(pose_dim * 4, 32), nn.BatchNorm1d(32), nn.ReLU(), nn.Linear(32, 32))
        self.gru = nn.GRU(self.in_size, hidden_size=self.hidden_size, num_layers=4, batch_first=True, bidirectional=True, dropout=0.3)
        self.out = nn.Sequential(nn.Linear(self.hidden_size, self.hidden_size // 2), nn.LeakyReLU(True),
                                 nn.Linear(self.hidden_size // 2, pose_dim))

    def forward(self, latent_code, pre_poses):
        raise RuntimeError('PoseDecoderGRU holds parameters only; run EmbeddingNet.forward')


class ContextEncoder(nn.Module):
    """embedding_net.py:220-259 (parameter container)."""

    def __init__(self, args, n_frames, n_words, word_embed_size, word_embeddings):
        super().__init__()
        from model.multimodal_context_net import TextEncoderTCN, WavEncoder       # deferred: that module imports this one (as in the reference)
        self.text_encoder = TextEncoderTCN(args, n_words, word_embed_size, pre_trained_embedding=word_embeddings)
        self.audio_encoder = WavEncoder()
        self.gru = nn.GRU(32 + 32, hidden_size=256, num_layers=2, bidirectional=False, batch_first=True)
        self.out = nn.Sequential(nn.Linear(256, 128), nn.BatchNorm1d(128), nn.ReLU(inplace=True), nn.Linear(128, 32))
        self.fc_mu = nn.Linear(32, 32)
        self.fc_logvar = nn.Linear(32, 32)
        self.do_flatten_parameters = False

    def forward(self, in_text, in_spec):
        raise RuntimeError('ContextEncoder holds parameters only; run EmbeddingNet.forward')


class EmbeddingNet(nn.Module):
    def __init__(self, args, pose_dim, n_frames, n_words, word_embed_size, word_embeddings, mode):
        super().__init__()
        if mode != 'pose':
            self.context_encoder = ContextEncoder(args, n_frames, n_words, word_embed_size, word_embeddings)
            self.pose_encoder = PoseEncoderConv(n_frames, pose_dim)
            self.decoder = PoseDecoderGRU(n_frames, pose_dim)
        else:
            self.context_encoder = None
            self.pose_encoder = PoseEncoderConv(n_frames, pose_dim)
            self.decoder = PoseDecoderConv(n_frames, pose_dim)
        self.mode = mode
        self._engine = None
        self._train_engine = None
        self._joint_engine = None
        self._noise_seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)

    def joint_engine(self) -> JointEmbeddingEngine:
        if self._joint_engine is None:
            self._joint_engine = JointEmbeddingEngine(self)
        return self._joint_engine

    def _forward_joint(self, in_text, in_audio, pre_poses, poses, input_mode, variational_encoding):
        """embedding_net.py:276-308 for the joint-embedding model; returns detached tensors (training goes through train_iter_embed)."""
        assert not variational_encoding, 'the reference runs the joint-embedding model with variational_encoding=False (train_joint_embed.py:12-15,56)'
        ref = poses if poses is not None else pre_poses
        if not ref.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('EmbeddingNet runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        if input_mode == 'random':
            input_mode = 'speech' if random.random() > 0.5 else 'pose'                # embedding_net.py:295-296
        assert input_mode in ('speech', 'pose')
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            assert in_text is not None and in_audio is not None and poses is not None
            self._joint_params = [p for p in self.parameters() if p.requires_grad]
            c_feat, c_mu, c_lv, p_feat, p_mu, p_lv, out = _JointEmbeddingFn.apply(
                self, in_text.contiguous(), in_audio.detach().contiguous().float(), pre_poses.detach().contiguous().float(),
                poses.detach().contiguous().float(), input_mode, *self._joint_params)
            return c_feat, c_mu, c_lv, p_feat, p_mu, p_lv, out
        eng = self.joint_engine().ensure(ref.device)
        r = eng.forward(in_text, in_audio, pre_poses, poses, input_mode, self.training)
        cl = lambda t: None if t is None else t.clone()
        return cl(r['c_feat']), cl(r['c_mu']), cl(r['c_lv']), cl(r['p_mu']), cl(r['p_mu']), cl(r['p_lv']), cl(r['out'])

    def engine(self) -> EmbeddingEngine:
        if self._engine is None:
            self._engine = EmbeddingEngine(self)
        return self._engine

    def train_engine(self) -> AutoEncoderTrainEngine:
        if self._train_engine is None:
            self._train_engine = AutoEncoderTrainEngine(self)
        return self._train_engine

    def forward(self, in_text, in_audio, pre_poses, poses, input_mode=None, variational_encoding=False):
        _lib.require_cuda()
        if input_mode is None:
            assert self.mode is not None
            input_mode = self.mode
        if self.context_encoder is not None:
            return self._forward_joint(in_text, in_audio, pre_poses, poses, input_mode, variational_encoding)
        assert input_mode == 'pose', "EmbeddingNet(mode='pose') has no context encoder (embedding_net.py:270-273,282)"
        if not poses.is_cuda and not _lib.TRACE_ONLY:
            raise _lib.TgError('EmbeddingNet runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
        if self.training:
            # batch-statistics forward (updates the BatchNorm running statistics like the reference's train-mode forward)
            assert not variational_encoding, 'the reference trains the auto-encoder with variational_encoding=False (train_feature_extractor.py:58)'
            poses_c = poses.detach().contiguous().float()
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                self._ae_params = [p for p in self.parameters() if p.requires_grad]
                feat, mu, logvar, recon = _AutoEncoderFn.apply(self, poses_c, *self._ae_params)
                return None, None, None, feat, mu, logvar, recon
            teng = self.train_engine().ensure(poses.device)
            mu, logvar, recon = teng.forward(poses_c, training=True)
            return None, None, None, mu.clone(), mu.clone(), logvar.clone(), recon.clone()
        eng = self.engine().ensure(poses.device)
        poses_c = poses.detach().contiguous().float()
        eps = None
        if variational_encoding:
            eps = eng.ws.get('emb.eps', (poses.shape[0], 32))
            off = eng.ws.get('emb.off', (1,), torch.int64, zero=True)
            ops.philox_normal(eps, eps.numel(), int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF), off, 9)
            ops.increment_i64(off, 1)
        feat, mu, logvar, recon = eng.forward(poses_c, eps, variational_encoding, decode=True)
        return None, None, None, feat.clone(), mu.clone(), logvar.clone(), recon.clone()

    def freeze_pose_nets(self):
        for param in self.pose_encoder.parameters():
            param.requires_grad = False
        for param in self.decoder.parameters():
            param.requires_grad = False
