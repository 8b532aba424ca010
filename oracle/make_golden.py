"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

The reference is imported from /root/reference/scripts with the three import shims of
SURVEY.md section 8c (stub ``fasttext`` / ``umap`` modules, ``model.embedding_net`` imported
first).  Deterministic weights from oracle/synth.py are loaded into the reference modules
with strict=True (which also proves the key/shape inventory matches), noise is injected
through the reference's own seams (``model.embedding_net.reparameterize`` is looked up as a
module attribute at multimodal_context_net.py:131; ``torch.randperm`` at train_gan.py:62;
``nn.Dropout`` sub-modules are swapped for mask-multiplying modules; the GRU inter-layer
dropout, which cannot take a mask, is set to p=0), and the outputs are stored.

Large tensors (gradients, post-step weights) are stored as (l2 norm, sum, 64 strided
samples) so that fixtures stay small.  TEST INFRASTRUCTURE — never imported by the product.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

from . import synth
from . import trimodal_oracle as O

REF = '/root/reference/scripts'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def import_reference():
    for name in ('fasttext', 'umap'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model.embedding_net as ref_embed          # must come first (circular import)
    import model.multimodal_context_net as ref_net
    import train_eval.train_gan as ref_gan
    from model import vocab as ref_vocab
    return ref_embed, ref_net, ref_gan, ref_vocab


def digest(t: torch.Tensor, n: int = 64) -> np.ndarray:
    """[l2, sum, samples...] in float64."""
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return np.concatenate([[f.norm().item(), f.sum().item()], f[idx].numpy()])


class MaskDrop(torch.nn.Module):
    """Stands in for nn.Dropout inside a reference module instance: multiplies by queued masks."""
    def __init__(self):
        super().__init__()
        self.queue = []

    def forward(self, x):
        if not self.queue:
            return x
        return x * self.queue.pop(0)


def golden_cfg() -> O.HotPathConfig:
    return O.HotPathConfig(n_words=500, n_speakers=20)


def build_reference(cfg, ref_net, ref_vocab, dropout_prob):
    args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, input_context='both',
                              hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=dropout_prob,
                              freeze_wordembed=False, z_type='speaker', loss_warmup=cfg.loss_warmup,
                              loss_gan_weight=cfg.loss_gan_weight, loss_regression_weight=cfg.loss_regression_weight,
                              loss_kld_weight=cfg.loss_kld_weight, loss_reg_weight=cfg.loss_reg_weight)
    spk = ref_vocab.Vocab('vid', insert_default_tokens=False)
    while spk.n_words < cfg.n_speakers:
        spk.index_word(f'spk{spk.n_words}')
    G = ref_net.PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, None, z_obj=spk)
    D = ref_net.ConvDiscriminator(cfg.pose_dim)
    G.load_state_dict(synth.with_tcn_aliases(synth.generator_state_dict(cfg)), strict=True)
    D.load_state_dict(synth.discriminator_state_dict(cfg), strict=True)
    return args, G, D


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    ref_embed, ref_net, ref_gan, ref_vocab = import_reference()
    cfg = golden_cfg()
    B = 3
    inp = synth.make_inputs(cfg, B, seed=1)
    pre_seq = O.make_pre_seq(inp['target'], cfg.n_pre_poses)

    # ---------------- 1. eval-mode forwards -------------------------------------------------
    args, G, D = build_reference(cfg, ref_net, ref_vocab, cfg.dropout_prob)
    G.eval(); D.eval()
    eps = synth.make_noise(cfg, B, seed=1).eps[0]
    orig_reparam = ref_embed.reparameterize
    ref_embed.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
    with torch.no_grad():
        poses, z, mu, logvar = G(pre_seq, inp['in_text'], inp['in_audio'], inp['vid'])
        audio_feat = G.audio_encoder(inp['in_audio'])
        text_feat, _ = G.text_encoder(inp['in_text'])
        d_real = D(inp['target'])
        d_fake = D(poses)
    np.savez(os.path.join(OUT, 'forward_eval.npz'), poses=poses.numpy(), z=z.numpy(), mu=mu.numpy(),
             logvar=logvar.numpy(), audio_feat=audio_feat.numpy(), text_feat=text_feat.numpy(),
             d_real=d_real.numpy(), d_fake=d_fake.numpy())
    print('forward_eval: poses l2', poses.norm().item(), 'd_real', d_real.flatten().tolist())

    # ---------------- 2. full train_iter_gan, epoch 11 (post warm-up) and epoch 0 ------------
    for tag, epoch, use_masks in (('train_e11', 11, True), ('train_e0', 0, False)):
        args, G, D = build_reference(cfg, ref_net, ref_vocab, cfg.dropout_prob)
        G.train(); D.train()
        G.gru.dropout = 0.0          # inter-layer GRU dropout cannot take a mask
        D.gru.dropout = 0.0
        noise = synth.golden_noise(cfg, B, 2, use_masks)
        # swap nn.Dropout modules for mask queues (reference *instances*, not sources)
        drops = {'emb': MaskDrop()}
        G.text_encoder.drop = drops['emb']
        for i, blk in enumerate(G.text_encoder.tcn.network):
            for j, pos in ((1, 3), (2, 7)):
                md = MaskDrop()
                drops[f'tcn{i}_{j}'] = md
                blk.net[pos] = md
        n_g_fwd = 3 if epoch > cfg.loss_warmup else 2
        fwd_ids = [0, 1, 2] if epoch > cfg.loss_warmup else [1, 2]
        if use_masks:
            for fid in fwd_ids:
                for k, md in drops.items():
                    md.queue.append(noise.g_masks[fid][k])
        eps_queue = [noise.eps[i] for i in fwd_ids]
        ref_embed.reparameterize = lambda mu, logvar: mu + eps_queue.pop(0) * torch.exp(0.5 * logvar)
        orig_randperm = torch.randperm
        torch.randperm = lambda n, *a, **k: noise.perm.clone()
        g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
        d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
        ret = ref_gan.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'],
                                     G, D, g_opt, d_opt)
        torch.randperm = orig_randperm
        assert not eps_queue, 'reparameterize call count mismatch'
        store = {f'loss_{k}': np.float64(v) for k, v in ret.items()}
        for k, p in G.named_parameters():
            store['ggrad/' + k] = digest(p.grad if p.grad is not None else torch.zeros_like(p))
        for k, v in G.state_dict().items():
            store['gpost/' + k] = digest(v)
        for k, v in D.state_dict().items():
            store['dpost/' + k] = digest(v)
        np.savez(os.path.join(OUT, tag + '.npz'), **store)
        print(tag, {k: float(v) for k, v in ret.items()})
    ref_embed.reparameterize = orig_reparam

    # ---------------- 3. EmbeddingNet(mode='pose') + FGD ------------------------------------
    args_e = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.3,
                                freeze_wordembed=False)
    E = ref_embed.EmbeddingNet(args_e, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None, 'pose')
    E.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    E.eval()
    rng = np.random.Generator(np.random.PCG64(77))
    real = torch.from_numpy((0.5 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim))).astype(np.float32))
    fake = torch.from_numpy((1.0 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim)) + 0.3).astype(np.float32))
    with torch.no_grad():
        _, _, _, rf, _, _, rrec = E(None, None, None, real, 'pose', variational_encoding=False)
        _, _, _, ff, _, _, frec = E(None, None, None, fake, 'pose', variational_encoding=False)
    import model.embedding_space_evaluator as ref_eval
    from scipy import linalg as sl

    class _L:                                               # SciPy >= 1.18 dropped disp= (SURVEY 8c shim 3)
        @staticmethod
        def sqrtm(a, disp=True):
            r = sl.sqrtm(a)
            return r if disp else (r, 0.0)
    ref_eval.linalg = _L
    ev = ref_eval.EmbeddingSpaceEvaluator.__new__(ref_eval.EmbeddingSpaceEvaluator)
    ev.generated_feat_list = [ff.numpy()]
    ev.real_feat_list = [rf.numpy()]
    fgd, feat_dist = ev.get_scores()
    np.savez(os.path.join(OUT, 'embedding_fgd.npz'), real_feat=rf.numpy(), fake_feat=ff.numpy(),
             real_recon=digest(rrec), fake_recon=digest(frec), fgd=np.float64(fgd), feat_dist=np.float64(feat_dist))
    print('fgd', fgd, 'feat_dist', feat_dist)


if __name__ == '__main__':
    main()
