"""tests/golden/forward_variants.npz: eval-mode PoseGenerator forwards of the UNMODIFIED reference module built with the other
constructor variants the drop-in boundary admits (args.input_context in audio / text / none, z_obj a Vocab / any truthy value / None;
scripts/model/multimodal_context_net.py:65-93,110-160).  TEST INFRASTRUCTURE ONLY.   python -m oracle.make_golden_variants"""
import argparse
import os

import numpy as np
import torch

from . import synth
from . import trimodal_oracle as O
from .make_golden import OUT, golden_cfg, import_reference

VARIANTS = (('audio', 'speaker'), ('text', 'random'), ('none', None), ('both', 'random'), ('none', 'speaker'))


def main():
    ref_embed, ref_net, _, ref_vocab = import_reference()
    cfg = golden_cfg()
    B = 3
    inp = synth.make_inputs(cfg, B, seed=1)
    pre_seq = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    eps = synth.make_noise(cfg, B, seed=1).eps[0]
    store = {}
    for ctx, zm in VARIANTS:
        args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, input_context=ctx, hidden_size=cfg.hidden_size,
                                  n_layers=cfg.n_layers, dropout_prob=cfg.dropout_prob, freeze_wordembed=False)
        z_obj = None
        if zm == 'speaker':
            z_obj = ref_vocab.Vocab('vid', insert_default_tokens=False)
            while z_obj.n_words < cfg.n_speakers:
                z_obj.index_word(f'spk{z_obj.n_words}')
        elif zm == 'random':
            z_obj = 1                                                   # train.py:85
        G = ref_net.PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, None, z_obj=z_obj)
        G.load_state_dict(synth.with_tcn_aliases(synth.generator_state_dict_variant(cfg, ctx, zm)), strict=True)
        G.eval()
        orig_reparam, orig_randn = ref_embed.reparameterize, torch.randn
        ref_embed.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
        torch.randn = lambda *a, **k: eps.clone()
        try:
            with torch.no_grad():
                poses, z, mu, logvar = G(pre_seq, inp['in_text'], inp['in_audio'], inp['vid'] if zm == 'speaker' else None)
        finally:
            ref_embed.reparameterize, torch.randn = orig_reparam, orig_randn
        tag = f'{ctx}_{zm}'
        store[tag + '/poses'] = poses.numpy()
        if z is not None:
            store[tag + '/z'] = z.numpy()
        print(tag, float(poses.norm()))
    np.savez(os.path.join(OUT, 'forward_variants.npz'), **store)


if __name__ == '__main__':
    main()
