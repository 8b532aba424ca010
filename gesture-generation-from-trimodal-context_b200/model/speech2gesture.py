"""Speech2Gesture baseline (reference: scripts/model/speech2gesture.py): drop-in `Generator` / `Discriminator` with the reference's
constructor signatures, sub-module tree and `state_dict` keys.  The torch sub-modules are PARAMETER CONTAINERS (same shapes, same default
initialisation order); `forward` runs the launch plan of tgb200/s2g_engine.py over the C ABI (im2col + tcgen05 TF32 GEMM convolutions,
BatchNorm / LeakyReLU / U-Net plumbing kernels of csrc/s2g.cu).  CUDA tensors only."""
import torch
import torch.nn as nn


class Conv2d_tf(nn.Conv2d):
    """Conv2d with TensorFlow padding semantics (speech2gesture.py:9-52): `padding` is 'SAME' or 'VALID'; the arithmetic lives in the engine."""

    def __init__(self, *args, padding='SAME', **kwargs):
        super().__init__(*args, **kwargs)
        self.padding = padding

    def forward(self, input):
        raise RuntimeError('Conv2d_tf is a parameter container here; call the owning Generator / Discriminator')


class Conv1d_tf(nn.Conv1d):
    """Conv1d with TensorFlow padding semantics (speech2gesture.py:55-101)."""

    def __init__(self, *args, padding='SAME', **kwargs):
        super().__init__(*args, **kwargs)
        self.padding = padding

    def forward(self, input):
        raise RuntimeError('Conv1d_tf is a parameter container here; call the owning Generator / Discriminator')


def ConvNormRelu(in_channels, out_channels, type='1d', downsample=False, k=None, s=None, padding='SAME'):
    """speech2gesture.py:104-117: conv (k3 s1, or k4 s2 when downsampling) + BatchNorm + LeakyReLU(0.2)."""
    if k is None and s is None:
        k, s = (4, 2) if downsample else (3, 1)
    assert type in ('1d', '2d')
    conv, norm = (Conv1d_tf, nn.BatchNorm1d) if type == '1d' else (Conv2d_tf, nn.BatchNorm2d)
    return nn.Sequential(conv(in_channels, out_channels, kernel_size=k, stride=s, padding=padding), norm(out_channels), nn.LeakyReLU(0.2, True))


class UnetUp(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = ConvNormRelu(in_ch, out_ch)


class AudioEncoder(nn.Module):
    """speech2gesture.py:133-195 (parameter container)."""

    def __init__(self, n_frames):
        super().__init__()
        self.n_frames = n_frames
        chans = [(1, 64, False), (64, 64, True), (64, 128, False), (128, 128, True), (128, 256, False), (256, 256, True), (256, 256, False)]
        self.first_net = nn.Sequential(*[ConvNormRelu(i, o, '2d', d) for (i, o, d) in chans], ConvNormRelu(256, 256, '2d', False, padding='VALID'))
        self.make_1d = torch.nn.Upsample((n_frames, 1), mode='bilinear', align_corners=False)
        self.down1 = nn.Sequential(ConvNormRelu(256, 256, '1d', False), ConvNormRelu(256, 256, '1d', False))
        for i in range(2, 7):
            setattr(self, 'down%d' % i, ConvNormRelu(256, 256, '1d', True))
        for i in range(1, 6):
            setattr(self, 'up%d' % i, UnetUp(256, 256))


class Generator(nn.Module):
    """speech2gesture.py:198-229: `Generator(n_poses, pose_dim, n_pre_poses)`, `forward(in_spec [B,128,L], pre_poses [B,n_pre,D]) -> [B,n_poses,D]`."""

    def __init__(self, n_poses, pose_dim, n_pre_poses):
        super().__init__()
        self.gen_length = n_poses
        self.audio_encoder = AudioEncoder(n_poses)
        self.pre_pose_encoder = nn.Sequential(nn.Linear(n_pre_poses * pose_dim, 32), nn.BatchNorm1d(32), nn.ReLU(inplace=True), nn.Linear(32, 16))
        self.decoder = nn.Sequential(ConvNormRelu(256 + 16, 256), ConvNormRelu(256, 256), ConvNormRelu(256, 256), ConvNormRelu(256, 256))
        self.final_out = nn.Conv1d(256, pose_dim, 1, 1)
        self._engine = None

    def engine(self):
        if self._engine is None:
            from tgb200.s2g_engine import S2GGeneratorEngine
            self._engine = S2GGeneratorEngine(self)
        return self._engine

    def forward(self, in_spec, pre_poses):
        """Inference / evaluation call sites (train.py:277-279, synthesize.py:137-139): no autograd graph; training goes through
        train_eval.train_speech2gesture.train_iter_speech2gesture."""
        eng = self.engine().ensure(in_spec.device)
        return eng.forward(in_spec, pre_poses, self.training).clone()


class Discriminator(nn.Module):
    """speech2gesture.py:232-250: `forward(x [B,T,D])` differences the poses in time itself and returns [B,1,T'] scores."""

    def __init__(self, pose_dim):
        super().__init__()
        self.net = nn.Sequential(
            Conv1d_tf(pose_dim, 64, kernel_size=4, stride=2, padding='SAME'),
            nn.LeakyReLU(0.2, True),
            ConvNormRelu(64, 128, '1d', True),
            ConvNormRelu(128, 256, '1d', k=4, s=1),
            Conv1d_tf(256, 1, kernel_size=4, stride=1, padding='SAME'),
        )
        self._engine = None

    def engine(self):
        if self._engine is None:
            from tgb200.s2g_engine import S2GDiscriminatorEngine
            self._engine = S2GDiscriminatorEngine(self)
        return self._engine

    def forward(self, x):
        eng = self.engine().ensure(x.device)
        out = eng.forward(x.contiguous().float(), self.training)              # [B, T', 1] channels-last
        return out.view(x.shape[0], -1, 1).transpose(1, 2).clone()
