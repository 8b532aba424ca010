"""Parity checks of the auto-encoder trainer shared by the CPU suite (launch plan executed on tests/cabi_emulator.py) and
the GPU suite (same plan on libtg_b200.so): OUR train_feature_extractor.train_iter / train_joint_embed.{train_iter_embed, eval_embed}
and the train-mode EmbeddingNet forward against the reference-executed golden (tests/golden/ae_train.npz) and the fp64 oracle."""
import os

import numpy as np
import torch

from conftest import GOLDEN, rel_l2
from oracle import embed_train_oracle as EO
from oracle import synth
from oracle.make_golden import digest, golden_cfg
from test_oracle_ae_golden import digest_close, post_close


def build(device):
    from model.embedding_net import EmbeddingNet
    cfg = golden_cfg()
    net = EmbeddingNet(None, cfg.pose_dim, cfg.n_poses, None, None, None, 'pose')          # train_feature_extractor.py:133
    net.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    net = net.to(device)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4, betas=(0.5, 0.999))                  # :134
    return cfg, net, opt


def check_against_golden(g, tag, net, loss, step, lr, tol):
    """Step 1 is held to fp32 round-off.  The first Adam update moves EVERY element by ~lr*sign(g), including those whose gradient is
    round-off noise, so from step 2 on two correct fp32 implementations follow trajectories that differ by +-lr in a few weights:
    gradients are then compared at 5e-3 (the reference-vs-itself spread on another BLAS), losses at 10*tol."""
    ltol = tol if step == 1 else 10 * tol
    gtol = 5 * tol if step == 1 else 5e-3
    assert abs(loss - float(g[f'{tag}/loss'])) <= ltol * abs(float(g[f'{tag}/loss'])), (tag, loss, float(g[f'{tag}/loss']))
    sd = net.state_dict()
    for k, p in net.named_parameters():
        gr = p.grad.detach().cpu()
        if k in EO.ZERO_GRAD_PARAMS:
            assert gr.abs().max().item() < 1e-4, k
            continue
        try:
            digest_close(digest(gr), g[f'{tag}/grad/{k}'], gtol)
        except AssertionError as exc:
            raise AssertionError('%s grad %s: %s' % (tag, k, exc)) from None
    for k, v in sd.items():
        ref = g[f'{tag}/post/{k}']
        if k in EO.ZERO_GRAD_PARAMS or (k in EO.NOISY_RUNNING_MEANS and step > 1):
            continue
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(ref[2]) == step, k
        elif 'running' in k:
            digest_close(digest(v.cpu()), ref, 5 * tol)
        else:
            post_close(digest(v.cpu()), ref, lr, step)


def run_feature_extractor_two_steps(device, tol=2e-5):
    import train_feature_extractor as tfx
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    cfg, net, opt = build(device)
    net.train()
    init = {k: v.clone() for k, v in net.state_dict().items()}
    for step in (1, 2):
        ret = tfx.train_iter(None, 0, torch.from_numpy(g[f'fx{step}/target']).to(device), net, opt)
        assert set(ret) == {'loss'}
        check_against_golden(g, f'fx{step}', net, ret['loss'], step, float(g['lr']), tol)
    for k in EO.UNUSED_PARAMS:                       # no gradient in the reference -> Adam leaves them alone
        assert torch.equal(net.state_dict()[k].cpu(), init[k].cpu()), k
    st = opt.state_dict()['state']                   # the caller's torch optimiser reflects our flat Adam state
    assert len(st) == len(list(net.parameters())) and int(st[0]['step']) == 2


def run_train_iter_embed(device, tol=2e-5):
    from train_eval.train_joint_embed import train_iter_embed
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    cfg, net, opt = build(device)
    net.train()
    ret = train_iter_embed(None, 0, None, None, torch.from_numpy(g['fx1/target']).to(device), net, opt)
    check_against_golden(g, 'je', net, ret['loss'], 1, float(g['lr']), tol)


def run_forward_and_eval(device, tol=2e-5):
    from train_eval.train_joint_embed import eval_embed
    g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
    cfg, net, opt = build(device)
    tgt = torch.from_numpy(g['fx1/target']).to(device)
    net.train()
    out = net(None, None, None, tgt, None, variational_encoding=False)
    assert out[0] is None and out[1] is None and out[2] is None
    assert rel_l2(out[3], g['fwd_train/feat']) < tol and rel_l2(out[4], g['fwd_train/feat']) < tol
    assert rel_l2(out[5], g['fwd_train/logvar']) < tol and rel_l2(out[6], g['fwd_train/recon']) < tol
    for k, v in net.state_dict().items():
        if 'running' in k:
            digest_close(digest(v.cpu()), g[f'fwd_train/post/{k}'], 5 * tol)
        elif 'num_batches' in k:
            assert int(v) == 1
    cfg, net, opt = build(device)
    net.eval()
    loss, recon = eval_embed(None, None, None, tgt, net)
    assert abs(float(loss) - float(g['eval/loss'])) < tol * float(g['eval/loss'])
    assert rel_l2(recon, g['eval/recon']) < tol
    # eval-mode train-engine forward == the FGD engine's eval forward (two launch plans, one network)
    out = net(None, None, None, tgt, 'pose', variational_encoding=False)
    assert rel_l2(out[6], g['eval/recon']) < tol


def run_full_batch_vs_fp64_oracle(device, B=128, steps=3, tol=1e-4):
    """BASELINE batch size: several consecutive steps (so CUDA-graph replay is exercised on the GPU) vs the float64 oracle."""
    import train_feature_extractor as tfx
    cfg, net, opt = build(device)
    net.train()
    sd = synth.embedding_net_state_dict(cfg)
    o_opt = synth.zeros_like_opt(sd)
    for step in range(1, steps + 1):
        tgt = synth.make_inputs(cfg, B, seed=30 + step)['target']
        want = EO.train_iter_ae_oracle(sd, o_opt, step, tgt, 5e-4, True, dtype=torch.float64)
        ret = tfx.train_iter(None, 0, tgt.to(device), net, opt)
        assert abs(ret['loss'] - want['loss']) <= tol * abs(want['loss']), (step, ret['loss'], want['loss'])
        worst = 0.0
        for k, p in net.named_parameters():
            if k in EO.ZERO_GRAD_PARAMS or k in EO.UNUSED_PARAMS:
                continue
            worst = max(worst, rel_l2(p.grad, want['grads'][k]))
        # from the second step on the two trajectories differ by the +-lr walk of the zero-gradient biases (absorbed by BatchNorm); the walk
        # follows the sign of round-off gradients, which split-K atomics reorder from run to run (seen: 0.6e-2 .. 2.3e-2 at step 4)
        assert worst < (tol if step == 1 else 500 * tol), (step, worst)
        sd, o_opt = {k: (v.float() if v.is_floating_point() else v) for k, v in want['sd'].items()}, want['opt']
    return worst
