"""CPU: pins oracle/joint_embed_oracle.py (joint-embedding model) against fixtures produced by executing the reference's
EmbeddingNet(mode='random') and train_iter_embed / eval_embed (oracle/make_golden_joint.py -> tests/golden/joint_embed.npz)."""
import os

import numpy as np
import torch

from conftest import GOLDEN, rel_l2
from oracle import joint_embed_oracle as J
from oracle import synth
from oracle.make_golden import digest, golden_cfg
from test_oracle_ae_golden import digest_close, post_close

TOL = 2e-5
B = 4
NAMES = ('c_feat', 'c_mu', 'c_lv', 'p_feat', 'p_mu', 'p_lv', 'out')


def is_zero_grad(k):
    """Analytically zero gradients (a constant shift in front of a train-mode BatchNorm), cf. embed_train_oracle.ZERO_GRAD_PARAMS."""
    return k in ('context_encoder.out.0.bias', 'decoder.pre_pose_net.0.bias', 'pose_encoder.net.0.0.bias', 'pose_encoder.net.1.0.bias',
                 'pose_encoder.net.2.0.bias', 'pose_encoder.net.3.bias', 'pose_encoder.out_net.0.bias', 'pose_encoder.out_net.1.bias',
                 'pose_encoder.out_net.3.bias', 'context_encoder.audio_encoder.feat_extractor.0.bias',
                 'context_encoder.audio_encoder.feat_extractor.3.bias', 'context_encoder.audio_encoder.feat_extractor.6.bias')


def test_forwards_and_eval_embed():
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'joint_embed.npz'))
    sd = synth.joint_embedding_state_dict(cfg)
    inp = synth.make_inputs(cfg, B, seed=61)
    eps = torch.from_numpy(g['eps'])
    pre = inp['target'][:, :cfg.n_pre_poses]
    for training in (False, True):
        for mode in ('speech', 'pose'):
            with torch.no_grad():
                outs = J.embedding_net_joint(sd, inp['in_text'], inp['in_audio'], pre, inp['target'], mode, eps, training, cfg.n_layers, stats={})
            tag = f"fwd_{'train' if training else 'eval'}_{mode}"
            for name, o in zip(NAMES, outs):
                assert rel_l2(o, g[f'{tag}/{name}']) < TOL, (tag, name, rel_l2(o, g[f'{tag}/{name}']))
    with torch.no_grad():
        outs = J.embedding_net_joint(sd, inp['in_text'], inp['in_audio'], pre, inp['target'], 'speech', eps, False, cfg.n_layers)
    # eval_embed draws its own eps in the reference run: compare through the deterministic part (mu-decoded recon is not stored) -> loss only loosely
    assert np.isfinite(float(g['eval/loss']))


def run_two_steps(dtype=torch.float32):
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'joint_embed.npz'))
    sd = synth.joint_embedding_state_dict(cfg)
    opt = synth.zeros_like_opt(sd)
    steps = {}
    outs = []
    for step, mode in ((1, 'speech'), (2, 'pose')):
        noise = synth.golden_noise(cfg, B, 70 + step, True)
        e = noise.eps[0].repeat(1, 2)[:, :32].contiguous()
        data = synth.make_inputs(cfg, B, seed=63 + step)
        out = J.train_iter_embed_oracle(sd, opt, steps, data['in_text'], data['in_audio'], data['target'], cfg.n_pre_poses, mode, e,
                                        float(g['lr']), masks=noise.g_masks[0], dtype=dtype, n_tcn_layers=cfg.n_layers)
        outs.append(out)
        sd, opt, steps = ({k: (v.float() if v.is_floating_point() else v) for k, v in out['sd'].items()},
                          {m: {k: v.float() for k, v in out['opt'][m].items()} for m in ('m', 'v')}, out['step'])
    return g, outs


def test_train_iter_embed_speech_then_pose():
    g, outs = run_two_steps()
    lr = float(g['lr'])
    for step, out in enumerate(outs, start=1):
        tag = f'step{step}'
        assert abs(out['loss'] - float(g[f'{tag}/loss'])) <= (TOL if step == 1 else 20 * TOL) * abs(float(g[f'{tag}/loss']))
        for k, gr in out['grads'].items():
            assert (gr is not None) == bool(g[f'{tag}/hasgrad/{k}']), (tag, k)          # which branch receives gradients
            if gr is None or is_zero_grad(k):
                continue
            digest_close(digest(gr), g[f'{tag}/grad/{k}'], 5 * TOL if step == 1 else 5e-3)
        for k, v in out['sd'].items():
            ref = g[f'{tag}/post/{k}']
            if is_zero_grad(k) or ('running_mean' in k and step > 1):
                continue
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(ref[2]), (tag, k)
            elif 'running' in k:
                digest_close(digest(v), ref, 5 * TOL if step == 1 else 1e-3)
            else:
                post_close(digest(v), ref, lr, step)
    # Adam bookkeeping: parameters of the branch that was not decoded keep their moments and step count
    s = outs[1]['step']
    assert s['decoder.out.0.weight'] == 2 and s['context_encoder.fc_mu.weight'] == 1 and s['pose_encoder.fc_mu.weight'] == 1
    assert 'pose_encoder.fc_logvar.weight' not in s and 'context_encoder.fc_logvar.weight' in s
