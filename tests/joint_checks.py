"""Parity checks of the joint-embedding model shared by the CPU suite (launch plan on tests/cabi_emulator.py) and the GPU
suite: OUR EmbeddingNet(mode='random') forward and train_iter_embed / eval_embed against the reference-executed golden
(tests/golden/joint_embed.npz) and the fp64 oracle (oracle/joint_embed_oracle.py)."""
import argparse
import os

import numpy as np
import torch

from conftest import GOLDEN, rel_l2
from gpu_util import masks_to_ours
from oracle import joint_embed_oracle as J
from oracle import synth
from oracle.make_golden import digest, golden_cfg
from test_oracle_ae_golden import digest_close, post_close
from test_oracle_joint_golden import NAMES, is_zero_grad

B = 4


def build(device, cfg=None, seed=0):
    from model.embedding_net import EmbeddingNet
    cfg = cfg or golden_cfg()
    args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, hidden_size=cfg.hidden_size, n_layers=cfg.n_layers,
                              dropout_prob=cfg.dropout_prob, freeze_wordembed=False)
    net = EmbeddingNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None, 'random')           # train.py:60-62
    net.load_state_dict(synth.with_tcn_aliases(synth.joint_embedding_state_dict(cfg, seed)), strict=True)
    net = net.to(device)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4, betas=(0.5, 0.999))
    return cfg, args, net, opt


def run_forwards(device, tol=2e-5):
    g = np.load(os.path.join(GOLDEN, 'joint_embed.npz'))
    cfg = golden_cfg()
    inp = {k: v.to(device) for k, v in synth.make_inputs(cfg, B, seed=61).items()}
    eps = torch.from_numpy(g['eps']).to(device)
    pre = inp['target'][:, :cfg.n_pre_poses]
    for training in (False, True):
        for mode in ('speech', 'pose'):
            cfg, args, net, opt = build(device)
            net.train(training)
            net.joint_engine().noise = dict(eps=eps, masks={}, gru_masks=[None] * 4)          # golden: dropout off, injected eps
            outs = net(inp['in_text'], inp['in_audio'], pre, inp['target'], mode, variational_encoding=False)
            tag = f"fwd_{'train' if training else 'eval'}_{mode}"
            for name, o in zip(NAMES, outs):
                assert rel_l2(o, g[f'{tag}/{name}']) < tol, (tag, name, rel_l2(o, g[f'{tag}/{name}']))
    # eval_embed (train.py:270): the reference draws its own eps there, so compare with the golden forward decoded from the injected eps
    cfg, args, net, opt = build(device)
    net.eval()
    from train_eval.train_joint_embed import eval_embed
    net.joint_engine().noise = dict(eps=eps)
    loss, recon = eval_embed(inp['in_text'], inp['in_audio'], pre, inp['target'], net, mode='speech')
    assert rel_l2(recon, g['fwd_eval_speech/out']) < tol
    want = np.abs(g['fwd_eval_speech/out'] - inp['target'].cpu().numpy()).mean()
    assert abs(float(loss) - want) < max(1e-5, tol) * want


def run_two_steps(device, tol=2e-5):
    """mode='random' with the coin patched to 'speech' then 'pose' (as the golden): losses, gradients, which parameters moved."""
    import random
    from train_eval.train_joint_embed import train_iter_embed
    g = np.load(os.path.join(GOLDEN, 'joint_embed.npz'))
    cfg, args, net, opt = build(device)
    net.train()
    init = {k: v.clone() for k, v in net.state_dict().items()}
    coins = [0.9, 0.1]
    orig = random.random
    random.random = lambda: coins.pop(0)
    try:
        for step in (1, 2):
            noise = synth.golden_noise(cfg, B, 70 + step, True)
            e = noise.eps[0].repeat(1, 2)[:, :32].contiguous().to(device)
            net.joint_engine().noise = dict(eps=e, masks=masks_to_ours(noise.g_masks[0], device), gru_masks=[None] * 4)
            data = {k: v.to(device) for k, v in synth.make_inputs(cfg, B, seed=63 + step).items()}
            ret = train_iter_embed(args, 0, data['in_text'], data['in_audio'], data['target'], net, opt, mode='random')
            tag = f'step{step}'
            ref = float(g[f'{tag}/loss'])
            assert abs(ret['loss'] - ref) <= (tol if step == 1 else 20 * tol) * abs(ref), (tag, ret['loss'], ref)
            live = ('context_encoder.', 'decoder.') if step == 1 else ('pose_encoder.', 'decoder.')
            for k, p in net.named_parameters():
                if not k.startswith(live) or is_zero_grad(k) or not bool(g[f'{tag}/hasgrad/{k}']):
                    continue
                try:
                    digest_close(digest(p.grad.detach().cpu()), g[f'{tag}/grad/{k}'], 5 * tol if step == 1 else 5e-3)
                except AssertionError as exc:
                    raise AssertionError('%s grad %s: %s' % (tag, k, exc)) from None
            for k, v in net.state_dict().items():
                if '.tcn.network.' in k and ('.net.0.' in k or '.net.4.' in k):
                    continue
                refv = g[f'{tag}/post/{k}']
                if is_zero_grad(k) or ('running_mean' in k and step > 1):
                    continue
                if k.endswith('num_batches_tracked'):
                    assert int(v) == int(refv[2]), (tag, k)
                elif 'running' in k:
                    digest_close(digest(v.cpu()), refv, 5 * tol if step == 1 else 1e-3)
                else:
                    try:
                        post_close(digest(v.cpu()), refv, float(g['lr']), step)
                    except AssertionError as exc:
                        raise AssertionError('%s post %s: %s' % (tag, k, exc)) from None
    finally:
        random.random = orig
    assert not coins
    # Adam skipped what had no gradient: the pose encoder's fc_logvar never moves; per-parameter step counts follow the branches
    for k in ('pose_encoder.fc_logvar.weight', 'pose_encoder.fc_logvar.bias'):
        assert torch.equal(net.state_dict()[k].cpu(), init[k].cpu())
    st = opt.state_dict()['state']
    names = [k for k, _ in net.named_parameters()]
    steps = {names[i]: int(s['step']) for i, s in st.items()}
    assert steps['decoder.out.0.weight'] == 2 and steps['context_encoder.fc_mu.weight'] == 1 and steps['pose_encoder.fc_mu.weight'] == 1


def run_batch_vs_fp64_oracle(device, Bn=8, tol=1e-4, gtol=None):
    """One step per branch with EVERY dropout mask injected (incl. the decoder GRU's inter-layer masks, which the reference cannot
    take) vs the float64 oracle."""
    from train_eval.train_joint_embed import train_iter_embed
    worst = 0.0
    for branch in ('speech', 'pose'):
        cfg, args, net, opt = build(device)
        net.train()
        sd = synth.joint_embedding_state_dict(cfg)
        data = synth.make_inputs(cfg, Bn, seed=80)
        noise = synth.make_noise(cfg, Bn, seed=81, dropout=True)
        e = noise.eps[0].repeat(1, 2)[:, :32].contiguous()
        gm = [noise.g_masks[0][f'gru{l}'] for l in range(3)] + [None]
        text_masks = {k: v for k, v in noise.g_masks[0].items() if not k.startswith('gru')}
        want = J.train_iter_embed_oracle(sd, synth.zeros_like_opt(sd), {}, data['in_text'], data['in_audio'], data['target'], cfg.n_pre_poses, branch,
                                         e, 5e-4, masks=text_masks, gru_masks=gm, dtype=torch.float64, n_tcn_layers=cfg.n_layers)
        net.joint_engine().noise = dict(eps=e.to(device), masks=masks_to_ours(text_masks, device),
                                        gru_masks=[None if m is None else m.reshape(-1, m.shape[-1]).contiguous().to(device) for m in gm])
        ret = train_iter_embed(args, 0, data['in_text'].to(device), data['in_audio'].to(device), data['target'].to(device), net, opt, mode=branch)
        assert abs(ret['loss'] - want['loss']) <= tol * abs(want['loss']), (branch, ret['loss'], want['loss'])
        for k, p in net.named_parameters():
            r = want['grads'][k]
            if r is None or is_zero_grad(k) or float(r.norm()) < 1e-7:
                continue
            err = rel_l2(p.grad, r)
            assert err < (gtol if gtol is not None else 10 * tol), (branch, k, err)
            worst = max(worst, err)
    return worst


def run_evaluate_testset(device):
    """train.py:234-329 for args.model 'joint_embedding' / 'gesture_autoencoder': metrics from the device accumulators vs the oracle."""
    import ae_checks
    from model.embedding_net import EmbeddingNet
    from model.embedding_space_evaluator import EmbeddingSpaceEvaluator
    from oracle import embed_train_oracle as EO
    from train_eval.evaluate import evaluate_testset
    cfg, args, net, opt = build(device)
    args.model = 'joint_embedding'
    batches = []
    for i in range(2):
        inp = synth.make_inputs(cfg, 6, seed=95 + i)
        batches.append((None, None, inp['in_text'], None, inp['target'], inp['in_audio'], None, None))
    enet = EmbeddingNet(None, cfg.pose_dim, cfg.n_poses, None, None, None, 'pose')
    enet.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    ev = EmbeddingSpaceEvaluator.from_net(enet, cfg.n_pre_poses, torch.device(device))
    net.train()
    ret = evaluate_testset(batches, net, None, ev, args)
    assert set(ret) == {'loss', 'joint_mae', 'frechet', 'feat_dist'} and net.training and ev.get_no_of_samples() == 2
    assert all(np.isfinite(v) for v in ret.values()) and ret['loss'] > 0 and ret['joint_mae'] > 0
    # the auto-encoder: loss only
    cfg2, aenet, _ = ae_checks.build(device)
    a2 = argparse.Namespace(model='gesture_autoencoder', n_pre_poses=cfg.n_pre_poses)
    ret = evaluate_testset(batches, aenet, None, None, a2)
    sd = synth.embedding_net_state_dict(cfg)
    want = np.mean([EO.eval_embed_oracle(sd, b[4])[0] for b in batches])
    assert set(ret) == {'loss', 'joint_mae'} and ret['joint_mae'] == 0.0
    assert abs(ret['loss'] - want) < 2e-5 * want, (ret['loss'], want)
