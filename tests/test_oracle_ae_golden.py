"""CPU: pins oracle/embed_train_oracle.py (FGD auto-encoder trainer) against fixtures produced by EXECUTING the
reference's own train_feature_extractor.train_iter / train_joint_embed.{train_iter_embed, eval_embed}
(oracle/make_golden_ae.py -> tests/golden/ae_train.npz)."""
import os

import numpy as np
import torch

from oracle import embed_train_oracle as E
from oracle import synth
from oracle.make_golden import digest, golden_cfg
from conftest import GOLDEN, rel_l2

TOL = 2e-5
ZERO_GRAD_KEYS = E.ZERO_GRAD_PARAMS
NOISY_STATS = E.NOISY_RUNNING_MEANS


def digest_close(a, ref, tol):
    a = np.asarray(a); ref = np.asarray(ref)
    scale = max(abs(ref[0]), 1e-12)
    if scale < 1e-4:
        assert abs(a[0]) < 1e-4
        return
    assert abs(a[0] - ref[0]) <= tol * scale + 1e-9, (a[0], ref[0])
    n = max(len(ref) - 2, 1)
    assert np.abs(a[2:] - ref[2:]).max() <= tol * 50 * scale / np.sqrt(n) + 5e-7


def post_close(a, ref, lr, step):
    """Post-Adam weights: an element whose gradient is round-off noise may land 2*lr away per step; the bulk must agree tightly."""
    d = np.abs(np.asarray(a)[2:] - np.asarray(ref)[2:])
    assert d.max() <= 2.2 * lr * step + 1e-6, d.max()
    assert np.median(d) <= 2e-6 + 1e-5 * np.abs(np.asarray(ref)[2:]).max(), np.median(d)


def check_step(g, tag, out, lr, step):
    assert abs(out['loss'] - float(g[f'{tag}/loss'])) <= TOL * abs(float(g[f'{tag}/loss']))
    for k, gr in out['grads'].items():
        if k in ZERO_GRAD_KEYS:
            assert gr.abs().max().item() < 1e-4
            continue
        digest_close(digest(gr), g[f'{tag}/grad/{k}'], 5 * TOL)
    for k, v in out['sd'].items():
        ref = g[f'{tag}/post/{k}']
        if k in ZERO_GRAD_KEYS or (k in NOISY_STATS and step > 1):
            continue
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(ref[2]) == step
        elif 'running' in k:
            digest_close(digest(v), ref, 5 * TOL)
        else:
            post_close(digest(v), ref, lr, step)


def test_train_iter_feature_extractor_two_steps():
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    lr = float(g['lr'])
    sd = synth.embedding_net_state_dict(cfg)
    opt = synth.zeros_like_opt(sd)
    for step in (1, 2):
        out = E.train_iter_ae_oracle(sd, opt, step, torch.from_numpy(g[f'fx{step}/target']), lr, use_diff=True)
        check_step(g, f'fx{step}', out, lr, step)
        sd, opt = out['sd'], out['opt']
    # parameters without a gradient are untouched (torch.optim.Adam skips them)
    init = synth.embedding_net_state_dict(cfg)
    for k in E.UNUSED_PARAMS:
        assert torch.equal(sd[k], init[k])


def test_train_iter_embed_pose_mode():
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    sd = synth.embedding_net_state_dict(cfg)
    opt = synth.zeros_like_opt(sd)
    out = E.train_iter_ae_oracle(sd, opt, 1, torch.from_numpy(g['fx1/target']), float(g['lr']), use_diff=False)
    check_step(g, 'je', out, float(g['lr']), 1)


def test_train_forward_and_eval_embed():
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    sd = synth.embedding_net_state_dict(cfg)
    tgt = torch.from_numpy(g['fx1/target'])
    stats = {}
    with torch.no_grad():
        feat, mu, logvar, recon = E.embedding_net_pose(sd, tgt, True, stats)
    assert rel_l2(feat, g['fwd_train/feat']) < TOL and rel_l2(logvar, g['fwd_train/logvar']) < TOL
    assert rel_l2(recon, g['fwd_train/recon']) < TOL
    for k, v in stats.items():
        if not k.endswith('num_batches_tracked'):
            digest_close(digest(v), g[f'fwd_train/post/{k}'], 5 * TOL)
    loss, recon = E.eval_embed_oracle(sd, tgt)
    assert abs(loss - float(g['eval/loss'])) < TOL * float(g['eval/loss'])
    assert rel_l2(recon, g['eval/recon']) < TOL


def test_fp64_oracle_agrees_with_fp32():
    """Accuracy yard-stick: the same step in float64 (what the GPU parity tests at batch 128 compare against)."""
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    sd = synth.embedding_net_state_dict(cfg)
    opt = synth.zeros_like_opt(sd)
    tgt = torch.from_numpy(g['fx1/target'])
    a = E.train_iter_ae_oracle(sd, opt, 1, tgt, float(g['lr']), True)
    b = E.train_iter_ae_oracle(sd, opt, 1, tgt, float(g['lr']), True, dtype=torch.float64)
    assert abs(a['loss'] - b['loss']) < 1e-5 * abs(b['loss'])
    assert rel_l2(a['recon'], b['recon']) < 1e-5
    for k in a['grads']:
        if k not in ZERO_GRAD_KEYS and k not in E.UNUSED_PARAMS:
            assert rel_l2(a['grads'][k], b['grads'][k]) < 2e-4, k
