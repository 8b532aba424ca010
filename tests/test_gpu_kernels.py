"""GPU: every C-ABI kernel family against a plain torch fp32/fp64 computation of the same op (per-op localisation of
any parity failure).  Tolerances are fp32 round-off level."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from tgb200 import _lib
    _lib.load()
    return torch.device('cuda:0')


@pytest.fixture(autouse=True)
def strict_fp32():
    from tgb200 import config
    old = config.set_mode('fp32')
    yield
    config.set_mode(old)


def _rand(*shape, dev, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed + sum(shape))
    return (scale * torch.randn(*shape, generator=g)).to(dev)


@pytest.mark.parametrize('M,K,N', [(4352, 300, 150), (4352, 108, 1800), (100, 27, 16), (7, 16, 16), (4352, 600, 1800), (300, 150, 27), (128, 28, 1)])
def test_linear_fwd_bwd(dev, M, K, N):
    from tgb200 import ops
    x, w, b = _rand(M, K, dev=dev), _rand(N, K, dev=dev, seed=1, scale=K ** -0.5), _rand(N, dev=dev, seed=2)
    y = torch.empty(M, N, device=dev)
    ops.linear(x, w, b, y, M=M, K=K, N=N)
    ref = (x.double() @ w.double().t() + b.double())
    assert rel_l2(y, ref) < 2e-6
    dy = _rand(M, N, dev=dev, seed=3)
    dx = torch.empty(M, K, device=dev)
    ops.linear_dgrad(dy, w, dx, M=M, K=K, N=N)
    assert rel_l2(dx, dy.double() @ w.double()) < 2e-6
    dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    ops.linear_wgrad(x, dy, dw, db, M=M, K=K, N=N)
    assert rel_l2(dw, dy.double().t() @ x.double()) < 5e-6
    assert rel_l2(db, dy.double().sum(0)) < 5e-6


@pytest.mark.parametrize('B,Tin,Cin,N,k,stride,dil,pad', [
    (3, 1313, 16, 32, 15, 6, 1, 0), (2, 217, 64, 32, 15, 6, 1, 0), (4, 34, 27, 16, 3, 1, 1, 0), (3, 30, 64, 64, 4, 2, 1, 0),
    (2, 500, 1, 16, 15, 5, 1, 160)])
def test_conv1d_fwd_dgrad_wgrad(dev, B, Tin, Cin, N, k, stride, dil, pad):
    from tgb200 import ops
    x = _rand(B, Tin, Cin, dev=dev)
    w = _rand(N, Cin, k, dev=dev, seed=1, scale=(Cin * k) ** -0.5)
    b = _rand(N, dev=dev, seed=2)
    Tout = (Tin + 2 * pad - dil * (k - 1) - 1) // stride + 1
    y = torch.empty(B, Tout, N, device=dev)
    ops.conv1d(x, w, b, y, B=B, Tin=Tin, Cin=Cin, N=N, k=k, stride=stride, dil=dil, pad=pad)
    xr = x.double().transpose(1, 2).requires_grad_(True)
    wr = w.double().requires_grad_(True)
    br = b.double().requires_grad_(True)
    ref = F.conv1d(xr, wr, br, stride=stride, padding=pad, dilation=dil)
    assert ref.shape[2] == Tout
    assert rel_l2(y, ref.transpose(1, 2)) < 2e-6
    dy = _rand(B, Tout, N, dev=dev, seed=3)
    ref.backward(dy.double().transpose(1, 2))
    dx = torch.full((B, Tin, Cin), float('nan'), device=dev)
    ops.conv1d_dgrad(dy, w, dx, B=B, Tin=Tin, Tout=Tout, Cin=Cin, N=N, k=k, stride=stride, dil=dil, pad=pad)
    assert rel_l2(dx, xr.grad.transpose(1, 2)) < 2e-6
    dw = torch.zeros_like(w); db = torch.zeros_like(b)
    ops.conv1d_wgrad(x, dy, dw, db, B=B, Tin=Tin, Tout=Tout, Cin=Cin, N=N, k=k, stride=stride, dil=dil, pad=pad)
    assert rel_l2(dw, wr.grad) < 5e-6
    assert rel_l2(db, br.grad) < 5e-6
    if Cin == 1:
        y2 = torch.empty_like(y)
        ops.conv1_direct(x.view(B, Tin), w, b, y2, B=B, Tin=Tin, Tout=Tout, N=N, taps=k, stride=stride, pad=pad)
        assert rel_l2(y2, ref.transpose(1, 2)) < 2e-6


@pytest.mark.parametrize('d', [1, 2, 4, 8])
def test_causal_dilated_conv(dev, d):
    """tcn.py:19-31: pad both sides + chomp == left pad only; fused bias+ReLU+dropout-mask epilogue."""
    from tgb200 import ops
    B, T, C, k = 5, 34, 300, 2
    x = _rand(B, T, C, dev=dev); w = _rand(C, C, k, dev=dev, seed=1, scale=(C * k) ** -0.5); b = _rand(C, dev=dev, seed=2)
    mask = (torch.rand(B * T, C, device=dev) > 0.3).float() / 0.7
    y = torch.empty(B, T, C, device=dev)
    ops.conv1d(x, w, b, y, B=B, Tin=T, Tout=T, Cin=C, N=C, k=k, dil=d, pad=(k - 1) * d, act1=ops.ACT_RELU, mask=mask)
    xr = x.double().transpose(1, 2).requires_grad_(True); wr = w.double().requires_grad_(True)
    c = F.conv1d(xr, wr, b.double(), padding=(k - 1) * d, dilation=d)[:, :, :-(k - 1) * d]
    ref = torch.relu(c) * mask.double().view(B, T, C).transpose(1, 2)
    assert rel_l2(y, ref.transpose(1, 2)) < 2e-6
    dy = _rand(B, T, C, dev=dev, seed=3)
    ref.backward(dy.double().transpose(1, 2))
    dc = torch.empty(B * T, C, device=dev)
    ops.relu_mask_bwd(dy, y, mask, dc, B * T * C)
    dx = torch.empty(B, T, C, device=dev)
    ops.conv1d_dgrad(dc, w, dx, B=B, Tin=T, Tout=T, Cin=C, N=C, k=k, dil=d, pad=(k - 1) * d)
    assert rel_l2(dx, xr.grad.transpose(1, 2)) < 3e-6
    dw = torch.zeros_like(w)
    ops.conv1d_wgrad(x, dc, dw, None, B=B, Tin=T, Tout=T, Cin=C, N=C, k=k, dil=d, pad=(k - 1) * d)
    assert rel_l2(dw, wr.grad) < 5e-6


def test_batchnorm_train_fwd_bwd(dev):
    from tgb200 import ops
    M, C = 3 * 1313, 32
    x = _rand(M, C, dev=dev) * 2 + 0.5
    gamma, beta = _rand(C, dev=dev, seed=1) * 0.1 + 1, _rand(C, dev=dev, seed=2) * 0.1
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    mean, rstd, scale, shift = (torch.empty(C, device=dev) for _ in range(4))
    ops.col_stats(x, C, M, C, sums)
    ops.bn_finalize(sums, M, C, 1e-5, 0.1, 3, gamma, beta, rm, rv, nbt, mean, rstd, scale, shift)
    y = torch.empty_like(x)
    ops.affine_lrelu(x, y, M, C, scale, shift, 0.3)
    bn = torch.nn.BatchNorm1d(C).to(dev).double()
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
    xr = x.double().requires_grad_(True)
    for _ in range(3):
        ref = F.leaky_relu(bn(xr), 0.3)
    assert rel_l2(y, ref) < 2e-6
    assert rel_l2(rm, bn.running_mean) < 1e-5 and rel_l2(rv, bn.running_var) < 1e-5 and int(nbt) == 3
    dy = _rand(M, C, dev=dev, seed=5)
    ref.backward(dy.double())
    bs = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    ops.bn_bwd_reduce(dy, x, M, C, mean, rstd, scale, shift, 0.3, bs)
    dx = torch.empty_like(x); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    ops.bn_bwd_apply(dy, x, dx, M, C, mean, rstd, scale, shift, 0.3, gamma, bs, dg, db)
    assert rel_l2(dx, xr.grad) < 1e-5
    assert rel_l2(dg, bn.weight.grad) < 1e-5 and rel_l2(db, bn.bias.grad) < 1e-5


def _gru_params(I, H, dev, seed=0):
    s = H ** -0.5
    return dict(wih=[_rand(3 * H, I, dev=dev, seed=seed + d, scale=s) for d in (0, 1)], whh=[_rand(3 * H, H, dev=dev, seed=seed + 2 + d, scale=s) for d in (0, 1)],
                bih=[_rand(3 * H, dev=dev, seed=seed + 4 + d, scale=s) for d in (0, 1)], bhh=[_rand(3 * H, dev=dev, seed=seed + 6 + d, scale=s) for d in (0, 1)])


@pytest.mark.parametrize('B,T,I,H', [(3, 34, 108, 300), (128, 34, 108, 300), (384, 34, 40, 300), (128, 28, 8, 64), (5, 7, 16, 200)])
def test_gru_layer_fwd_bwd(dev, B, T, I, H):
    """Persistent recurrence kernels vs the oracle cell (trimodal_oracle.gru_cell_sequence) in float64 + autograd."""
    from oracle import trimodal_oracle as O
    from tgb200 import ops
    p = _gru_params(I, H, dev)
    x = _rand(B, T, I, dev=dev)
    M = B * T
    wih = torch.cat(p['wih'], 0).contiguous(); bih = torch.cat(p['bih'], 0).contiguous()
    gi = torch.empty(M, 6 * H, device=dev)
    ops.linear(x.view(M, I), wih, bih, gi, M=M, K=I, N=6 * H)
    whhT = [torch.empty(H, 3 * H, device=dev) for _ in (0, 1)]
    for d in (0, 1):
        ops.transpose(p['whh'][d], whhT[d], 3 * H, H)
        assert torch.equal(whhT[d], p['whh'][d].t().contiguous())
    out = torch.full((M, 2 * H), float('nan'), device=dev)
    saved = torch.empty(4, M, 2 * H, device=dev)
    sync = torch.zeros(max(ops.gru_sync_ints(B, H), 1), dtype=torch.int32, device=dev)
    ops.gru_layer_fwd(gi, whhT[0], whhT[1], p['bhh'][0], p['bhh'][1], out, saved, M * 2 * H, sync, B, T, H)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    xd = x.double().requires_grad_(True)
    pd = {k: [t.double().requires_grad_(True) for t in v] for k, v in p.items()}
    ref = torch.cat([O.gru_cell_sequence(xd, pd['wih'][d], pd['whh'][d], pd['bih'][d], pd['bhh'][d], bool(d)) for d in (0, 1)], dim=2)
    assert rel_l2(out, ref) < 5e-6, rel_l2(out, ref)
    # backward on a batch slice of the forward (what train_iter_gan does)
    lo, hi = (0, B) if B < 8 else (B // 4, B // 4 + max(B // 2, 1))
    Bb = hi - lo
    dout = _rand(B, T, 2 * H, dev=dev, seed=9)
    dsel = torch.zeros_like(dout); dsel[lo:hi] = dout[lo:hi]
    ref.backward(dsel.double())
    Mb = Bb * T
    dgi = torch.full((Mb, 6 * H), float('nan'), device=dev); dgh = torch.full((Mb, 6 * H), float('nan'), device=dev)
    partial = torch.empty(max(ops.gru_bwd_scratch_floats(Bb, H), 1), device=dev)
    bsync = torch.zeros(max(ops.gru_sync_ints(Bb, H), 1), dtype=torch.int32, device=dev)
    ops.gru_layer_bwd(dout[lo:hi].contiguous().view(Mb, 2 * H), out[lo * T:hi * T], saved[0, lo * T:hi * T], M * 2 * H, p['whh'][0], p['whh'][1],
                      dgi, dgh, partial, bsync, Bb, T, H)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    # input grad and weight grads from dgi / dgh
    dx = torch.empty(Mb, I, device=dev)
    ops.linear_dgrad(dgi, wih, dx, M=Mb, K=I, N=6 * H)
    assert rel_l2(dx, xd.grad[lo:hi].reshape(Mb, I)) < 2e-5, rel_l2(dx, xd.grad[lo:hi].reshape(Mb, I))
    dwih = torch.zeros_like(wih); dbih = torch.zeros_like(bih)
    ops.linear_wgrad(x[lo:hi].reshape(Mb, I).contiguous(), dgi, dwih, dbih, M=Mb, K=I, N=6 * H)
    assert rel_l2(dwih, torch.cat([pd['wih'][0].grad, pd['wih'][1].grad], 0)) < 2e-5
    assert rel_l2(dbih, torch.cat([pd['bih'][0].grad, pd['bih'][1].grad], 0)) < 2e-5
    o_sl = out[lo * T:hi * T]
    for d in (0, 1):
        dwhh = torch.zeros(3 * H, H, device=dev); dbhh = torch.zeros(3 * H, device=dev)
        ops.conv_wgrad(o_sl[:, d * H:], dgh[:, d * 3 * H:], dwhh, B=Bb, Tin=T, Tout=T, N=3 * H, Cin=H, taps=1, pad=(1 if d == 0 else -1),
                       lda=2 * H, ldg=6 * H, ldw=H, dbias=dbhh)
        assert rel_l2(dwhh, pd['whh'][d].grad) < 2e-5, (d, rel_l2(dwhh, pd['whh'][d].grad))
        assert rel_l2(dbhh, pd['bhh'][d].grad) < 2e-5


def test_embedding_weightnorm_misc(dev):
    from oracle import trimodal_oracle as O
    from tgb200 import ops
    V, E, M = 500, 300, 3 * 34
    table = _rand(V, E, dev=dev)
    idx = torch.randint(0, V, (M,), device=dev)
    mask = (torch.rand(2 * M, E, device=dev) > 0.1).float() / 0.9
    out = torch.empty(2 * M, E, device=dev)
    ops.embedding_gather(table, idx, M, mask, out, 2 * M, E)
    assert torch.allclose(out, table[idx.repeat(2)] * mask)
    dt = torch.zeros_like(table)
    dout = _rand(M, E, dev=dev, seed=4)
    ops.embedding_scatter_add(dout, idx, mask[:M], dt, M, E)
    ref = torch.zeros_like(table).index_add_(0, idx, dout * mask[:M])
    assert rel_l2(dt, ref) < 1e-6
    N, K = 300, 600
    v = _rand(N, K, dev=dev, seed=5).requires_grad_(True); g = (_rand(N, dev=dev, seed=6).abs() + 0.5).requires_grad_(True)
    Cin, k = K // 2, 2
    w = torch.empty(k, N, Cin, device=dev); wT = torch.empty(k, Cin, N, device=dev); inv = torch.empty(N, device=dev)
    ops.weight_norm_fwd(v.data, g.data, w, wT, inv, N, Cin, k)
    ref = O.weight_norm_weight(g.view(N, 1, 1), v.view(N, Cin, k))           # [N, Cin, k]
    assert rel_l2(w, ref.permute(2, 0, 1)) < 1e-6                            # tap-major [k][N][Cin]
    assert rel_l2(wT, ref.permute(2, 1, 0)) < 1e-6
    dw = _rand(k, N, Cin, dev=dev, seed=7)
    ref.backward(dw.permute(1, 2, 0))
    dv = torch.zeros(N, K, device=dev); dg = torch.zeros(N, device=dev)
    ops.weight_norm_bwd(dw, v.data, g.data, inv, dv, dg, N, Cin, k)
    assert rel_l2(dv, v.grad) < 1e-5 and rel_l2(dg, g.grad) < 1e-5


def test_losses_and_adam(dev):
    from oracle import trimodal_oracle as O
    from tgb200 import ops
    cfg = O.HotPathConfig()
    B, T, D, Z = 16, 34, 27, 16
    out = (_rand(B, T, D, dev=dev) * 0.2).requires_grad_(True); tgt = _rand(B, T, D, dev=dev, seed=1) * 0.2
    outr = _rand(B, T, D, dev=dev, seed=2) * 0.2
    z, zr = _rand(B, Z, dev=dev, seed=3), _rand(B, Z, dev=dev, seed=4)
    zr[0] = z[0] + 1e-9                      # exercises the clamp(min=-1000) branch
    mu = _rand(B, Z, dev=dev, seed=5).requires_grad_(True); lv = (_rand(B, Z, dev=dev, seed=6) * 0.3).requires_grad_(True)
    dprob = torch.sigmoid(_rand(B, 1, dev=dev, seed=7)).requires_grad_(True)
    loss, hub, gen, div, kld = O.gen_losses(cfg, out, tgt, dprob, outr, z, zr, mu, lv, True)
    loss.backward()
    sc = torch.zeros(8, dtype=torch.float64, device=dev)
    d_out = torch.empty(B, T, D, device=dev); dmu = torch.empty(B, Z, device=dev); dlv = torch.empty(B, Z, device=dev)
    ops.gen_losses(out.data, tgt, outr, z, zr, mu.data, lv.data, B, T * D, Z, cfg.loss_regression_weight, cfg.loss_reg_weight, cfg.loss_kld_weight,
                   sc, d_out, dmu, dlv)
    dlogit = torch.empty(B, 1, device=dev)
    ops.bce_sigmoid(dprob.data, B, 1.0, 0.0, cfg.loss_gan_weight, sc[3:], dlogit)
    s = sc.cpu().tolist()
    assert abs(s[0] / (B * T * D) - hub.item()) < 1e-5 * abs(hub.item())
    assert abs(s[1] / B - div.item()) < 1e-5 * abs(div.item())
    assert abs(-0.5 * s[2] / (B * Z) - kld.item()) < 1e-5 * abs(kld.item())
    assert abs(s[3] - gen.item()) < 1e-5 * abs(gen.item())
    assert rel_l2(d_out, out.grad) < 1e-5
    assert rel_l2(dmu, mu.grad) < 1e-5 and rel_l2(dlv, lv.grad) < 1e-5
    assert rel_l2(dlogit, dprob.grad * dprob.data * (1 - dprob.data)) < 1e-5
    # dis loss pieces
    pr, pf = torch.sigmoid(_rand(B, 1, dev=dev, seed=8)), torch.sigmoid(_rand(B, 1, dev=dev, seed=9))
    sc.zero_()
    ops.bce_sigmoid(pr, B, 1.0, 0.0, 1.0, sc[4:], dlogit)
    ops.bce_sigmoid(pf, B, -1.0, 1.0, 1.0, sc[5:], dlogit)
    s = sc.cpu().tolist()
    assert abs(s[4] + s[5] - O.dis_loss(pr, pf).item()) < 1e-5
    # Adam, 3 steps, odd length (tail path)
    n = 1003 * 4 + 3
    p = _rand(n, dev=dev, seed=11); p0 = p.clone()
    m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    pm, pv, pp = torch.zeros(n, device=dev, dtype=torch.float64), torch.zeros(n, device=dev, dtype=torch.float64), p0.double()
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    pad = torch.zeros(((n + 3) // 4) * 4, device=dev)
    for it in range(3):
        g = _rand(n, dev=dev, seed=20 + it) * 0.01
        ops.increment_i64(step, 1)
        ops.adam_flat(p, g, m, v, n, 5e-4, 0.5, 0.999, 1e-8, 1.0, step)
        pp, pm, pv = O.adam_step(pp, g.double(), pm, pv, it + 1, 5e-4)
    assert rel_l2(p - p0, pp - p0.double()) < 1e-4


def test_philox_rng_statistics(dev):
    from tgb200 import ops
    n = 1 << 20
    off = torch.zeros(1, dtype=torch.int64, device=dev)
    x = torch.empty(n, device=dev)
    ops.philox_normal(x, n, 1234, off, 0)
    assert abs(x.mean().item()) < 5e-3 and abs(x.std().item() - 1) < 5e-3
    assert abs((x ** 4).mean().item() - 3.0) < 0.1
    m = torch.empty(n + 1, device=dev)[1:]          # unaligned tail path
    ops.philox_dropout_mask(m, n, 0.3, 1234, off, 1)
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.7) < 3e-3
    assert torch.allclose(m[m > 0], torch.tensor(1 / 0.7, device=dev))
    m2 = torch.empty(n, device=dev)
    ops.increment_i64(off, 1)
    ops.philox_dropout_mask(m2, n, 0.3, 1234, off, 1)
    assert (m2 != m).float().mean().item() > 0.3          # a new offset gives a new mask
    perm = torch.empty(128, dtype=torch.int64, device=dev)
    ops.philox_randperm(perm, 128, 99, off, 2)
    assert sorted(perm.cpu().tolist()) == list(range(128)) and perm.cpu().tolist() != list(range(128))


def test_feature_stats(dev):
    from tgb200 import ops
    n, Fd = 1000, 32
    x = _rand(n, Fd, dev=dev) * 2 + 1
    acc = torch.zeros(1 + Fd + Fd * Fd, dtype=torch.float64, device=dev)
    ops.feature_stats(x, n, Fd, acc)
    a = acc.cpu().numpy()
    xs = x.double().cpu().numpy()
    assert a[0] == n
    np.testing.assert_allclose(a[1:1 + Fd], xs.sum(0), rtol=1e-10)
    np.testing.assert_allclose(a[1 + Fd:].reshape(Fd, Fd), xs.T @ xs, rtol=1e-9)
