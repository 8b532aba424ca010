"""GPU: the FGD auto-encoder trainer through libtg_b200.so - the same checks the CPU suite runs on the emulated
launch plan (tests/ae_checks.py), plus unit tests of the two kernels it adds and of the transposed-convolution operator forms
against float64 torch.  (File name sorts last on purpose: the suite runs with -x and this is the newest path.)"""
import pytest
import torch
import torch.nn.functional as F

import ae_checks
from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from tgb200 import _lib
    _lib.load()
    return torch.device('cuda:0')


def _rand(*shape, dev, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed + sum(shape))
    return (scale * torch.randn(*shape, generator=g)).to(dev)


@pytest.mark.parametrize('B,R,C', [(128, 12, 32), (128, 4, 34), (3, 34, 4), (1, 1, 7), (257, 33, 5)])
def test_transpose_batched(dev, B, R, C):
    from tgb200 import ops
    x = _rand(B, R, C, dev=dev)
    out = torch.empty(B, C, R, device=dev)
    ops.transpose_batched(x, out, B, R, C)
    assert torch.equal(out, x.transpose(1, 2).contiguous())


@pytest.mark.parametrize('B,T,D,use_diff,weight', [(128, 34, 27, True, 1.0), (5, 34, 27, False, 1.0), (3, 2, 5, True, 100.0)])
def test_ae_recon_loss_value_and_gradient(dev, B, T, D, use_diff, weight):
    from tgb200 import ops
    r = _rand(B, T, D, dev=dev).double().requires_grad_(True)
    y = _rand(B, T, D, dev=dev, seed=1).double()
    y[0, 0, 0] = r.detach()[0, 0, 0]                                   # an exact tie: d|x|/dx = 0 there, like torch
    l0 = (r - y).abs().mean(dim=(1, 2))
    loss = l0.clone()
    if use_diff and T > 1:
        loss = loss + ((r[:, 1:] - r[:, :-1]) - (y[:, 1:] - y[:, :-1])).abs().mean(dim=(1, 2))
    (weight * loss.sum()).backward()
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    d = torch.empty(B, T, D, device=dev)
    ops.ae_recon_loss(r.detach().float(), y.float(), B, T, D, use_diff, weight, acc, d)
    assert abs(acc[0].item() - loss.sum().item()) < 1e-5 * abs(loss.sum().item())
    assert abs(acc[1].item() - l0.sum().item()) < 1e-5 * abs(l0.sum().item())
    assert rel_l2(d, r.grad) < 1e-6
    acc.zero_()
    ops.ae_recon_loss(r.detach().float(), y.float(), B, T, D, use_diff, weight, acc, None)      # value only (eval_embed)
    assert abs(acc[1].item() - l0.sum().item()) < 1e-5 * abs(l0.sum().item())


@pytest.mark.parametrize('B,L,ci,co,slope', [(128, 34, 4, 32, None), (128, 36, 32, 32, 0.2), (3, 5, 8, 16, 0.2)])
def test_conv_transpose_forms(dev, B, L, ci, co, slope):
    """ConvTranspose1d(k=3) forward = conv1d_dgrad, data gradient = conv1d, weight gradient = conv_wgrad with dilation -1
    (+ bias gradient), optionally with the BatchNorm+LeakyReLU prologue on the input - vs float64 F.conv_transpose1d autograd."""
    from tgb200 import ops
    k = 3
    x = _rand(B, L, ci, dev=dev)
    w = _rand(ci, co, k, dev=dev, seed=1, scale=(ci * k) ** -0.5)
    b = _rand(co, dev=dev, seed=2)
    sc, sh = (_rand(ci, dev=dev, seed=3).abs() + 0.5, _rand(ci, dev=dev, seed=4)) if slope is not None else (None, None)
    pro = dict(pscale=sc, pshift=sh, pslope=slope) if slope is not None else {}
    xd = x.double().requires_grad_(True); wd = w.double().requires_grad_(True); bd = b.double().requires_grad_(True)
    a = xd if slope is None else F.leaky_relu(xd * sc.double() + sh.double(), slope)
    ref = F.conv_transpose1d(a.transpose(1, 2), wd, bd).transpose(1, 2)                       # [B, L+2, co]
    dy = _rand(B, L + k - 1, co, dev=dev, seed=5)
    ref.backward(dy.double())
    y = torch.empty(B * (L + k - 1), co, device=dev)
    ops.conv1d_dgrad(x, w, y, B=B, Tin=L + k - 1, Tout=L, Cin=co, N=ci, k=k, bias=b, **pro)
    assert rel_l2(y, ref) < 2e-6
    dw = torch.zeros_like(w); db = torch.zeros_like(b)
    ops.conv_wgrad(x, dy, dw, B=B, Tin=L, Tout=L + k - 1, N=co, Cin=ci, taps=k, stride=1, dil=-1, pad=0, ldw=k, wsj=1, wsc=co * k, dbias=db, **pro)
    assert rel_l2(dw, wd.grad) < 5e-6 and rel_l2(db, bd.grad) < 5e-6
    if slope is None:
        dx = torch.empty(B * L, ci, device=dev)
        ops.conv1d(dy, w, None, dx, B=B, Tin=L + k - 1, Cin=co, N=ci, k=k)
        assert rel_l2(dx, xd.grad) < 2e-6


def test_feature_extractor_train_iter_vs_reference_golden(dev):
    ae_checks.run_feature_extractor_two_steps(dev)


def test_train_iter_embed_vs_reference_golden(dev):
    ae_checks.run_train_iter_embed(dev)


def test_train_forward_and_eval_embed_vs_reference_golden(dev):
    ae_checks.run_forward_and_eval(dev)


def test_batch128_steps_vs_fp64_oracle_with_graph_replay(dev):
    """Four consecutive batch-128 steps: two eager, then the captured CUDA graph (capture + replay, then a second replay), each vs the
    float64 oracle (step 1 at 1e-4; later steps only loosely: the trajectories separate by the +-lr walk of round-off gradients)."""
    from tgb200 import config
    assert config.graphs()
    ae_checks.run_full_batch_vs_fp64_oracle(dev, B=128, steps=4)


def test_graph_replay_matches_eager(dev):
    """Same data, same initial weights: 4 steps with CUDA-graph replay == 4 eager steps (post-step weights bit-for-bit or round-off)."""
    import train_feature_extractor as tfx
    from oracle import synth
    from tgb200 import config
    finals = []
    for graphs in (True, False):
        old = config.set_graphs(graphs)
        try:
            cfg, net, opt = ae_checks.build(dev)
            net.train()
            losses = [tfx.train_iter(None, 0, synth.make_inputs(cfg, 128, seed=40 + s)['target'].to(dev), net, opt)['loss'] for s in range(4)]
            finals.append((losses, {k: v.clone() for k, v in net.state_dict().items()}))
        finally:
            config.set_graphs(old)
    (la, a), (lb, b) = finals
    assert max(abs(x - y) / abs(y) for x, y in zip(la, lb)) < 1e-5
    from oracle import embed_train_oracle as EO
    for k in a:
        if k in EO.ZERO_GRAD_PARAMS or k in EO.NOISY_RUNNING_MEANS or not a[k].is_floating_point():
            continue
        if 'running' in k:            # statistics of activations whose weights differ by a few lr: compare relative to their size
            assert (a[k] - b[k]).abs().max().item() <= 2e-3 * max(1.0, b[k].abs().max().item()), k
            continue
        assert (a[k] - b[k]).abs().max().item() <= 2.2 * 5e-4 * 4 + 1e-6, k       # atomics reorder: a sign(g) flip moves an element by 2*lr
        assert (a[k] - b[k]).abs().median().item() < 2e-5, k


def test_evaluate_testset_and_cpu_refusal(dev):
    import train_feature_extractor as tfx
    from oracle import embed_train_oracle as EO
    from oracle import synth
    from tgb200 import _lib
    cfg, net, opt = ae_checks.build(dev)
    batches = [synth.make_inputs(cfg, n, seed=50 + i)['target'] for i, n in enumerate((8, 8, 4))]
    ret = tfx.evaluate_testset([(None, b.to(dev)) for b in batches], net)
    assert net.training                                                   # train_feature_extractor.py:44 "back to training mode"
    sd = synth.embedding_net_state_dict(cfg)
    want = sum(EO.eval_embed_oracle(sd, b)[0] * b.shape[0] for b in batches) / sum(b.shape[0] for b in batches)
    assert abs(ret['loss'] - want) < 2e-5 * want
    with pytest.raises(_lib.TgError):
        tfx.train_iter(None, 0, batches[0], net, opt)                     # CPU tensor: no fallback
