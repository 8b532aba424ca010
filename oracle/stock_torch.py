"""Stock-PyTorch baseline of the hot path.  TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

What a user of the reference gets on a B200 today: the same network expressed with torch.nn layers, i.e. cuDNN convolutions,
the cuDNN RNN for nn.GRU and cuBLAS for the linears, driven by autograd and torch.optim.Adam.  /root/reference does not exist
on the GPU box, so this file restates the module graph and the adversarial iteration from the cited lines; ``state_dict`` keys equal the
reference's, so the deterministic weights of oracle/synth.py load with strict=True and tests/test_stock_torch_pinned.py pins this
file to the oracle (which is itself pinned to the executed reference, tests/test_oracle_golden.py).

bench.py times it as ``gpu_stock_baseline`` (fp32 and TF32-allowed): the "existing Blackwell kernels" bar of SURVEY.md 2.3 / 8d.

Reference anchors: scripts/model/multimodal_context_net.py:9-28 (WavEncoder), :31-61 + scripts/model/tcn.py:7-64 (TextEncoderTCN),
:64-160 (PoseGenerator), :207-252 (ConvDiscriminator), scripts/model/embedding_net.py:10-13 (reparameterize),
scripts/train_eval/train_gan.py:13-103 (train_iter_gan), scripts/train.py:104-109 (optimisers)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import weight_norm


def _bn_lrelu(c, slope):
    return [nn.BatchNorm1d(c), nn.LeakyReLU(slope, inplace=True)]


class StockWavEncoder(nn.Module):          # multimodal_context_net.py:9-28
    def __init__(self):
        super().__init__()
        layers = [nn.Conv1d(1, 16, 15, stride=5, padding=1600)] + _bn_lrelu(16, 0.3)
        layers += [nn.Conv1d(16, 32, 15, stride=6)] + _bn_lrelu(32, 0.3)
        layers += [nn.Conv1d(32, 64, 15, stride=6)] + _bn_lrelu(64, 0.3)
        layers += [nn.Conv1d(64, 32, 15, stride=6)]
        self.feat_extractor = nn.Sequential(*layers)

    def forward(self, wav):
        return self.feat_extractor(wav.unsqueeze(1)).transpose(1, 2)


class _CausalBlock(nn.Module):             # tcn.py:17-46 (pad both sides by (k-1)*d, chomp the right side)
    def __init__(self, c_in, c_out, k, dilation, p):
        super().__init__()
        self.trim = (k - 1) * dilation
        self.conv1 = weight_norm(nn.Conv1d(c_in, c_out, k, padding=self.trim, dilation=dilation))
        self.conv2 = weight_norm(nn.Conv1d(c_out, c_out, k, padding=self.trim, dilation=dilation))
        self.p = p
        self.downsample = nn.Conv1d(c_in, c_out, 1) if c_in != c_out else None

    def forward(self, x):
        y = x
        for conv in (self.conv1, self.conv2):
            y = F.dropout(F.relu(conv(y)[:, :, :-self.trim].contiguous()), self.p, self.training)
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _Tcn(nn.Module):                     # tcn.py:49-64
    def __init__(self, c_in, channels, k, p):
        super().__init__()
        self.network = nn.Sequential(*[_CausalBlock(c_in if i == 0 else channels[i - 1], c, k, 2 ** i, p) for i, c in enumerate(channels)])

    def forward(self, x):
        return self.network(x)


class StockTextEncoder(nn.Module):         # multimodal_context_net.py:31-61
    def __init__(self, n_words, embed=300, hidden=300, n_layers=4, k=2, dropout=0.3, emb_dropout=0.1):
        super().__init__()
        self.embedding = nn.Embedding(n_words, embed)
        self.drop = nn.Dropout(emb_dropout)
        self.tcn = _Tcn(embed, [hidden] * n_layers, k, dropout)
        self.decoder = nn.Linear(hidden, 32)

    def forward(self, ids):
        y = self.tcn(self.drop(self.embedding(ids)).transpose(1, 2)).transpose(1, 2)
        return self.decoder(y)


class StockPoseGenerator(nn.Module):       # multimodal_context_net.py:64-160 (input_context='both', speaker-embedding z)
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.audio_encoder = StockWavEncoder()
        self.text_encoder = StockTextEncoder(cfg.n_words, cfg.wordembed_dim, cfg.hidden_size, cfg.n_layers, 2, cfg.dropout_prob, cfg.emb_dropout)
        self.speaker_embedding = nn.Sequential(nn.Embedding(cfg.n_speakers, cfg.z_size), nn.Linear(cfg.z_size, cfg.z_size))
        self.speaker_mu = nn.Linear(cfg.z_size, cfg.z_size)
        self.speaker_logvar = nn.Linear(cfg.z_size, cfg.z_size)
        self.gru = nn.GRU(cfg.gru_in, cfg.hidden_size, num_layers=cfg.n_layers, batch_first=True, bidirectional=True, dropout=cfg.dropout_prob)
        self.out = nn.Sequential(nn.Linear(cfg.hidden_size, cfg.hidden_size // 2), nn.LeakyReLU(True), nn.Linear(cfg.hidden_size // 2, cfg.pose_dim))

    def forward(self, pre_seq, in_text, in_audio, vid, eps=None):
        h = self.speaker_embedding(vid)
        mu, logvar = self.speaker_mu(h), self.speaker_logvar(h)
        eps = torch.randn_like(mu) if eps is None else eps              # embedding_net.py:10-13
        z = mu + eps * torch.exp(0.5 * logvar)
        x = torch.cat((pre_seq, self.audio_encoder(in_audio), self.text_encoder(in_text), z.unsqueeze(1).expand(-1, pre_seq.shape[1], -1)), dim=2)
        y, _ = self.gru(x)
        H = self.cfg.hidden_size
        y = y[..., :H] + y[..., H:]
        return self.out(y.reshape(-1, H)).reshape(x.shape[0], x.shape[1], -1), z, mu, logvar


class StockConvDiscriminator(nn.Module):   # multimodal_context_net.py:207-252
    def __init__(self, cfg):
        super().__init__()
        self.hidden = cfg.d_hidden
        self.pre_conv = nn.Sequential(nn.Conv1d(cfg.pose_dim, 16, 3), nn.BatchNorm1d(16), nn.LeakyReLU(True),
                                      nn.Conv1d(16, 8, 3), nn.BatchNorm1d(8), nn.LeakyReLU(True), nn.Conv1d(8, 8, 3))
        self.gru = nn.GRU(8, cfg.d_hidden, num_layers=cfg.d_layers, bidirectional=True, dropout=0.3, batch_first=True)
        self.out = nn.Linear(cfg.d_hidden, 1)
        self.out2 = nn.Linear(cfg.n_poses - 6, 1)

    def forward(self, poses):
        f = self.pre_conv(poses.transpose(1, 2)).transpose(1, 2)
        y, _ = self.gru(f)
        y = y[..., :self.hidden] + y[..., self.hidden:]
        y = self.out(y.reshape(-1, self.hidden)).view(poses.shape[0], -1)
        return torch.sigmoid(self.out2(y))


def build(cfg, g_sd, d_sd, device):
    """Modules with the given reference-keyed weights (strict load) and the reference's two Adam optimisers (train.py:104-109)."""
    G, D = StockPoseGenerator(cfg), StockConvDiscriminator(cfg)
    G.load_state_dict({k.replace('text_encoder.dropout', 'text_encoder.drop'): v for k, v in g_sd.items()}, strict=True)
    D.load_state_dict(d_sd, strict=True)
    G, D = G.to(device).train(), D.to(device).train()
    g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
    return G, D, g_opt, d_opt


def train_iter_gan_stock(cfg, epoch, in_text, in_audio, target, vid, G, D, g_opt, d_opt, eps=None, perm=None):
    """train_gan.py:13-103 with autograd and torch.optim (eps: optional 3 reparameterisation draws, perm: optional speaker permutation)."""
    eps = [None, None, None] if eps is None else eps
    pre_seq = target.new_zeros(target.shape[0], target.shape[1], target.shape[2] + 1)
    pre_seq[:, :cfg.n_pre_poses, :-1] = target[:, :cfg.n_pre_poses]
    pre_seq[:, :cfg.n_pre_poses, -1] = 1
    adversarial = epoch > cfg.loss_warmup and cfg.loss_gan_weight > 0.0
    dis = None
    if adversarial:
        d_opt.zero_grad()
        fake = G(pre_seq, in_text, in_audio, vid, eps[0])[0].detach()
        dis = -(torch.log(D(target) + 1e-8) + torch.log(1 - D(fake) + 1e-8)).mean()
        dis.backward()
        d_opt.step()
    g_opt.zero_grad()
    out, z, mu, logvar = G(pre_seq, in_text, in_audio, vid, eps[1])
    huber = F.smooth_l1_loss(out / 0.1, target / 0.1) * 0.1
    gen = -torch.log(D(out) + 1e-8).mean()
    perm = torch.randperm(vid.shape[0], device=vid.device) if perm is None else perm
    out_r, z_r, _, _ = G(pre_seq, in_text, in_audio, vid[perm], eps[2])
    pose_l1 = (F.smooth_l1_loss(out / 0.05, out_r.detach() / 0.05, reduction='none') * 0.05).sum(dim=(1, 2))
    z_l1 = (z.detach() - z_r.detach()).abs().mean(dim=1)
    div = torch.clamp(-(pose_l1 / (z_l1 + 1.0e-5)), min=-1000).mean()
    kld = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
    loss = cfg.loss_regression_weight * huber + cfg.loss_kld_weight * kld + cfg.loss_reg_weight * div
    if epoch > cfg.loss_warmup:
        loss = loss + cfg.loss_gan_weight * gen
    loss.backward()
    g_opt.step()
    ret = {'loss': cfg.loss_regression_weight * huber.item(), 'KLD': cfg.loss_kld_weight * kld.item(), 'DIV_REG': cfg.loss_reg_weight * div.item()}
    if adversarial:
        ret['gen'] = cfg.loss_gan_weight * gen.item()
        ret['dis'] = dis.item()
    return ret
