"""GPU: tcgen05 / TMA / TMEM kernels of the fast (TF32) mode against float64 torch.  TF32 keeps a 10-bit mantissa, so a
K-long dot product is accurate to ~1e-3 relative; the stated fast-mode tolerance on poses/losses is 1e-2 (north_star)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TF32_TOL = 2e-3


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _rand(*shape, dev, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed + sum(shape))
    return (scale * torch.randn(*shape, generator=g)).to(dev)


@pytest.mark.parametrize('M,N,K', [(128, 32, 32), (4352, 1800, 600), (4352, 1800, 108), (13056, 300, 300), (102, 150, 300), (4352, 32, 300),
                                   (1000, 64, 960), (4352, 600, 1800), (300, 28, 64),
                                   (13056, 1800, 600), (300, 1900, 64), (256, 960, 96),       # 240-column tiles
                                   (300, 320, 1024), (4352, 600, 1824)])                     # 256-row split-K tiles (with (4352, 600, 1800) above)
def test_gemm_tf32_plain(dev, M, N, K):
    from tgb200 import ops
    a = _rand(M, K, dev=dev); w = _rand(N, K, dev=dev, seed=1, scale=K ** -0.5); b = _rand(N, dev=dev, seed=2)
    c = torch.full((M, N), float('nan'), device=dev)
    ops.gemm_tf32(a, w, c, M=M, N=N, K=K, bias=b)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + b.double()
    assert rel_l2(c, ref) < TF32_TOL, rel_l2(c, ref)
    # epilogue: relu, mask, residual, second relu, accumulate
    mask = (torch.rand(M, N, device=dev) > 0.3).float() / 0.7
    res = _rand(M, N, dev=dev, seed=3)
    c2 = torch.ones(M, N, device=dev)
    ops.gemm_tf32(a, w, c2, M=M, N=N, K=K, bias=b, act1=ops.ACT_RELU, mask=mask, residual=res, act2=ops.ACT_RELU, accumulate=True)
    ref2 = torch.relu(torch.relu(ref) * mask.double() + res.double()) + 1.0
    assert rel_l2(c2, ref2) < TF32_TOL, rel_l2(c2, ref2)
    # linear epilogue (bias + mask): the shape class that runs split-K with red.global.add when K is long and the tiles are few
    c3 = torch.full((M, N), float('nan'), device=dev)
    ops.gemm_tf32(a, w, c3, M=M, N=N, K=K, bias=b, mask=mask)
    assert rel_l2(c3, ref * mask.double()) < TF32_TOL, rel_l2(c3, ref * mask.double())
    # the same with accumulate: split-K shapes add their partial tiles to the caller's C with red.global.add, no zero fill
    c4 = torch.full((M, N), 0.5, device=dev)
    ops.gemm_tf32(a, w, c4, M=M, N=N, K=K, bias=b, mask=mask, accumulate=True)
    assert rel_l2(c4, ref * mask.double() + 0.5) < TF32_TOL, rel_l2(c4, ref * mask.double() + 0.5)


@pytest.mark.parametrize('d', [1, 2, 4, 8])
@pytest.mark.parametrize('B', [3, 128, 384])
def test_gemm_tf32_causal_two_tap(dev, B, d):
    """TCN conv (k=2, dilation d, causal) and its anti-causal data gradient as two-tap GEMMs (csrc/gemm_tcn.cu: clip-group tiles, the
    shifted tap zero-filled per clip by the TMA unit).  B = 3: one ragged tile, N split over two CTAs; 128: 43 tiles, the last one with two
    of its three clips; 384: 128 tiles with the whole N = 300 in one 304-column accumulator (the forward sweep's shape)."""
    from tgb200 import ops
    T, C = 34, 300
    x = _rand(B, T, C, dev=dev); w = _rand(C, C, 2, dev=dev, seed=1, scale=(2 * C) ** -0.5); b = _rand(C, dev=dev, seed=2)
    wt = w.permute(2, 0, 1).contiguous()                       # [tap][N][Cin]
    y = torch.full((B * T, C), float('nan'), device=dev)
    ops.gemm_tf32(x.view(B * T, C), wt.view(2 * C, C), y, M=B * T, N=C, K=C, taps=2, shift0=-d, T=T, bias=b, act1=ops.ACT_RELU)
    xr = x.double().transpose(1, 2).requires_grad_(True)
    ref = torch.relu(F.conv1d(xr, w.double(), b.double(), padding=d, dilation=d)[:, :, :-d])
    assert rel_l2(y.view(B, T, C), ref.transpose(1, 2)) < TF32_TOL
    # TemporalBlock tail in the epilogue (tcn.py:28-29,46): xo = relu(relu(conv + b) * mask + residual), and its backward head
    mask = (torch.rand(B * T, C, device=dev) > 0.3).float() / 0.7
    res = _rand(B * T, C, dev=dev, seed=4)
    xo = torch.full((B * T, C), float('nan'), device=dev)
    ops.gemm_tf32(x.view(B * T, C), wt.view(2 * C, C), xo, M=B * T, N=C, K=C, taps=2, shift0=-d, T=T, bias=b, act1=ops.ACT_RELU, mask=mask,
                  residual=res, act2=ops.ACT_RELU)
    y2 = ref.detach().transpose(1, 2).reshape(B * T, C) * mask.double()
    assert rel_l2(xo, torch.relu(y2 + res.double())) < TF32_TOL
    g = _rand(B * T, C, dev=dev, seed=5)
    dpre = torch.full((B * T, C), float('nan'), device=dev); dc2 = torch.full((B * T, C), float('nan'), device=dev)
    ops.tcn_res_bwd(g, xo, res, mask, dpre, dc2, B * T * C)
    y2_gpu = (y * mask)                                     # what the un-fused plan would have stored
    assert torch.equal(dpre, g * (xo > 0))
    want = dpre * mask * (y2_gpu > 0)
    assert (dc2 != want).float().mean().item() < 1e-4        # identical except where y2 is absorbed by the rounding of y2 + x
    dy = _rand(B, T, C, dev=dev, seed=3)
    ref.backward(dy.double().transpose(1, 2))
    dc = (dy * (ref.transpose(1, 2) > 0)).float().contiguous()
    # dx[t] = dc[t+d] W0 + dc[t] W1 : B operand = W_tap^T  ([tap][Cin][N])
    wtt = w.permute(2, 1, 0).contiguous()
    dx = torch.full((B * T, C), float('nan'), device=dev)
    ops.gemm_tf32(dc.view(B * T, C), wtt.view(2 * C, C), dx, M=B * T, N=C, K=C, taps=2, shift0=d, T=T)
    assert rel_l2(dx.view(B, T, C), xr.grad.transpose(1, 2)) < 2 * TF32_TOL


@pytest.mark.parametrize('B,T,H', [(384, 34, 300), (5, 7, 200)])
def test_gru_layer_tensor_core_fwd_fused_dropout(dev, B, T, H):
    """tg_gru_layer_fwd_tf32_drop: same recurrence, and the masked copy of the output (nn.GRU's inter-layer dropout) written by the kernel."""
    from test_gpu_kernels import _gru_params
    from tgb200 import ops
    p = _gru_params(16, H, dev)
    M = B * T
    gi = _rand(M, 6 * H, dev=dev)
    mask = (torch.rand(M, 2 * H, generator=torch.Generator().manual_seed(5)) > 0.3).float().div(0.7).to(dev)
    sync = torch.zeros(max(ops.gru_tf32_sync_ints(B, H), 1), dtype=torch.int32, device=dev)
    outs = []
    for fused in (False, True):
        out = torch.full((M, 2 * H), float('nan'), device=dev); saved = torch.empty(4, M, 2 * H, device=dev)
        drop = torch.full((M, 2 * H), float('nan'), device=dev)
        if fused:
            ops.gru_layer_fwd_tf32_drop(gi, p['whh'][0], p['whh'][1], p['bhh'][0], p['bhh'][1], out, saved, M * 2 * H, mask, drop, sync, B, T, H)
        else:
            ops.gru_layer_fwd_tf32(gi, p['whh'][0], p['whh'][1], p['bhh'][0], p['bhh'][1], out, saved, M * 2 * H, sync, B, T, H)
        if dev.type == 'cuda':
            torch.cuda.synchronize()
        outs.append((out, saved, drop))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[1][2], outs[1][0] * mask)


@pytest.mark.parametrize('B,T,I,H', [(3, 34, 108, 300), (128, 34, 108, 300), (384, 34, 40, 300), (128, 28, 8, 64), (5, 7, 16, 200), (600, 5, 16, 300)])
def test_gru_layer_tensor_core_fwd_bwd(dev, B, T, I, H):
    """tcgen05 recurrence kernels (W_hh resident in smem, TMEM accumulators) vs the float64 oracle cell."""
    from oracle import trimodal_oracle as O
    from test_gpu_kernels import _gru_params
    from tgb200 import ops
    p = _gru_params(I, H, dev)
    x = _rand(B, T, I, dev=dev)
    M = B * T
    wih = torch.cat(p['wih'], 0).contiguous(); bih = torch.cat(p['bih'], 0).contiguous()
    gi = torch.empty(M, 6 * H, device=dev)
    ops.linear(x.view(M, I), wih, bih, gi, M=M, K=I, N=6 * H)
    out = torch.full((M, 2 * H), float('nan'), device=dev)
    saved = torch.empty(4, M, 2 * H, device=dev)
    sync = torch.zeros(max(ops.gru_tf32_sync_ints(B, H), 1), dtype=torch.int32, device=dev)
    ops.gru_layer_fwd_tf32(gi, p['whh'][0], p['whh'][1], p['bhh'][0], p['bhh'][1], out, saved, M * 2 * H, sync, B, T, H)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    xd = x.double().requires_grad_(True)
    pd = {k: [t.double().requires_grad_(True) for t in v] for k, v in p.items()}
    ref = torch.cat([O.gru_cell_sequence(xd, pd['wih'][d], pd['whh'][d], pd['bih'][d], pd['bhh'][d], bool(d)) for d in (0, 1)], dim=2)
    e = rel_l2(out, ref)
    assert e < TF32_TOL, e
    lo, hi = (0, B) if B < 8 else (B // 4, B // 4 + max(B // 3, 1))
    Bb = hi - lo
    dout = _rand(B, T, 2 * H, dev=dev, seed=9)
    dsel = torch.zeros_like(dout); dsel[lo:hi] = dout[lo:hi]
    ref.backward(dsel.double())
    Mb = Bb * T
    dgi = torch.full((Mb, 6 * H), float('nan'), device=dev); dgh = torch.full((Mb, 6 * H), float('nan'), device=dev)
    partial = torch.empty(max(ops.gru_bwd_tf32_scratch_floats(Bb, H), 1), device=dev)
    bsync = torch.zeros(max(ops.gru_tf32_sync_ints(Bb, H), 1), dtype=torch.int32, device=dev)
    whhT = [p['whh'][d].t().contiguous() for d in (0, 1)]
    ops.gru_layer_bwd_tf32(dout[lo:hi].contiguous().view(Mb, 2 * H), out[lo * T:hi * T], saved[0, lo * T:hi * T], M * 2 * H, whhT[0], whhT[1],
                           dgi, dgh, partial, bsync, Bb, T, H)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    dx = torch.empty(Mb, I, device=dev)
    ops.linear_dgrad(dgi, wih, dx, M=Mb, K=I, N=6 * H)
    e = rel_l2(dx, xd.grad[lo:hi].reshape(Mb, I))
    assert e < 3 * TF32_TOL, e
    o_sl = out[lo * T:hi * T]
    for d in (0, 1):
        dwhh = torch.zeros(3 * H, H, device=dev); dbhh = torch.zeros(3 * H, device=dev)
        ops.conv_wgrad(o_sl[:, d * H:], dgh[:, d * 3 * H:], dwhh, B=Bb, Tin=T, Tout=T, N=3 * H, Cin=H, taps=1, pad=(1 if d == 0 else -1),
                       lda=2 * H, ldg=6 * H, ldw=H, dbias=dbhh)
        e = rel_l2(dwhh, pd['whh'][d].grad)
        assert e < 3 * TF32_TOL, (d, e)
        assert rel_l2(dbhh, pd['bhh'][d].grad) < 3 * TF32_TOL


@pytest.mark.parametrize('B,T,N,Cin,shift', [(128, 34, 1800, 600, 0), (128, 34, 300, 300, -4), (128, 34, 900, 300, 1), (3, 34, 900, 300, -1),
                                              (128, 34, 150, 300, 0), (5, 34, 32, 300, 0), (128, 34, 300, 300, 8)])
def test_wgrad_tf32(dev, B, T, N, Cin, shift):
    """MN-major tcgen05 weight gradient: dW += G^T X with clip-local shifted rows of X (zero outside the clip)."""
    from tgb200 import ops
    M = B * T
    ldg = (N + 8 + 3) // 4 * 4                          # pitched views (16-byte aligned rows), like dgh[:, d*3H:] / out[:, d*H:]
    G = _rand(M, ldg, dev=dev)[:, :N]
    X = _rand(M, Cin + 4, dev=dev, seed=1)[:, :Cin]
    dW = torch.ones(N, Cin, device=dev)
    db = torch.zeros(N, device=dev)
    ops.wgrad_tf32(G, X, dW, B=B, T=T, N=N, Cin=Cin, shift=shift, ldg=ldg, ldx=Cin + 4, dbias=db)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    Xs = torch.zeros(B, T, Cin, device=dev, dtype=torch.float64)
    X3 = X.double().reshape(B, T, Cin)
    if shift == 0:
        Xs = X3
    elif shift > 0:
        Xs[:, :T - shift] = X3[:, shift:]
    else:
        Xs[:, -shift:] = X3[:, :T + shift]
    ref = G.double().t() @ Xs.reshape(M, Cin) + 1.0
    e = rel_l2(dW, ref)
    assert e < TF32_TOL, e
    assert rel_l2(db, G.double().sum(0)) < 1e-5


@pytest.mark.parametrize('B,Tin,cin,cout,k,s', [(3, 200, 16, 32, 15, 6), (128, 1313, 32, 64, 15, 6), (128, 217, 64, 32, 15, 6),
                                                (4, 7891, 16, 32, 15, 6)])
def test_strided_conv_window_gemm_fwd_bwd(dev, B, Tin, cin, cout, k, s):
    """WavEncoder conv2-4 (multimodal_context_net.py:16-22) on the tensor cores: forward = TF32 GEMM over the overlapping-window
    view (tg_gemm_tf32 clip mode), weight gradient = tg_wgrad_tf32 over the same view, data gradient = column GEMM + tg_col2im."""
    from tgb200 import ops
    Tout = (Tin - k) // s + 1
    x = _rand(B, Tin, cin, dev=dev); w = _rand(cout, cin, k, dev=dev, seed=1, scale=(cin * k) ** -0.5); b = _rand(cout, dev=dev, seed=2)
    w2 = torch.empty(cout, k * cin, device=dev); w2t = torch.empty(k * cin, cout, device=dev)
    ops.window_weights(w, w2, w2t, cout, cin, k)
    assert torch.equal(w2, w.permute(0, 2, 1).reshape(cout, k * cin)) and torch.equal(w2t, w2.t())
    y = torch.full((B * Tout, cout), float('nan'), device=dev)
    ops.gemm_tf32(x.view(B * Tin, cin), w2, y, M=B * Tout, N=cout, K=k * cin, lda=s * cin, clip_rows=Tout, a_clip_pitch=Tin * cin, bias=b)
    xr = x.double().transpose(1, 2).requires_grad_(True)
    wr = w.double().requires_grad_(True); br = b.double().requires_grad_(True)
    ref = F.conv1d(xr, wr, br, stride=s)
    assert rel_l2(y.view(B, Tout, cout), ref.transpose(1, 2)) < TF32_TOL
    dy = _rand(B, Tout, cout, dev=dev, seed=3)
    ref.backward(dy.double().transpose(1, 2))
    # weight gradient (accumulates)
    dw2 = torch.zeros(cout, k * cin, device=dev); db = torch.ones(cout, device=dev); dw = torch.ones(cout, cin, k, device=dev)
    ops.wgrad_tf32(dy.view(B * Tout, cout), x.view(B * Tin, cin), dw2, B=B, T=Tout, N=cout, Cin=k * cin, ldx=s * cin, x_clip_pitch=Tin * cin,
                   dbias=db)
    ops.window_wgrad_add(dw2, dw, cout, cin, k)
    assert rel_l2(dw - 1.0, wr.grad) < TF32_TOL, rel_l2(dw - 1.0, wr.grad)
    assert rel_l2(db - 1.0, br.grad) < 1e-5
    # data gradient
    col = torch.empty(B * Tout, k * cin, device=dev); da = torch.full((B * Tin, cin), float('nan'), device=dev)
    ops.gemm_tf32(dy.view(B * Tout, cout), w2t, col, M=B * Tout, N=k * cin, K=cout)
    ops.col2im(col, da, B=B, Tin=Tin, Tout=Tout, Cin=cin, k=k, stride=s)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    assert rel_l2(da.view(B, Tin, cin), xr.grad.transpose(1, 2)) < TF32_TOL


@pytest.mark.parametrize('B,Tin', [(2, 1000), (128, 36267), (5, 36266)])
def test_conv1_wgrad(dev, B, Tin):
    """Weight / bias gradient of WavEncoder conv1 (Conv1d(1,16,15,stride 5,pad 1600), multimodal_context_net.py:13), fp32."""
    from tgb200 import ops
    k, s, pad, N = 15, 5, 1600, 16
    Tout = (Tin + 2 * pad - k) // s + 1
    x = _rand(B, Tin, dev=dev); dy = _rand(B, Tout, N, dev=dev, seed=1)
    w = torch.zeros(N, 1, k, dtype=torch.float64, device=dev, requires_grad=True); bb = torch.zeros(N, dtype=torch.float64, device=dev, requires_grad=True)
    ref = F.conv1d(x.double().unsqueeze(1), w, bb, stride=s, padding=pad)
    ref.backward(dy.double().transpose(1, 2))
    dw = torch.ones(N, k, device=dev); db = torch.ones(N, device=dev)
    ops.conv1_wgrad(x, dy.view(B * Tout, N), dw, db, B=B, Tin=Tin, Tout=Tout, N=N, taps=k, stride=s, pad=pad)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    assert rel_l2(dw - 1.0, w.grad.view(N, k)) < 1e-4, rel_l2(dw - 1.0, w.grad.view(N, k))
    assert rel_l2(db - 1.0, bb.grad) < 1e-4


@pytest.mark.parametrize('B,Tin,cin,cout,k,s', [(3, 200, 16, 32, 15, 6), (128, 7891, 16, 32, 15, 6), (128, 1313, 32, 64, 15, 6), (128, 217, 64, 32, 15, 6),
                                                (5, 100, 8, 16, 5, 5), (2, 61, 4, 8, 7, 3)])
def test_conv_dgrad_tf32_matches_conv_transpose(dev, B, Tin, cin, cout, k, s):
    """tg_conv_dgrad_tf32 (accumulating shifted taps, no column matrix) against autograd's conv1d data gradient."""
    from tgb200 import ops
    Tout = (Tin - k) // s + 1
    w = _rand(cout, cin, k, dev=dev, scale=(cin * k) ** -0.5)
    dy = _rand(B, Tout, cout, dev=dev, seed=1)
    ntap = -(-k // s)
    wd = torch.full((ntap * s * cin, cout), float('nan'), device=dev)
    ops.window_dgrad_weights(w, wd, cout, cin, k, s)
    da = torch.full((B, Tin, cin), float('nan'), device=dev)
    guard = torch.full((64,), 7.0, device=dev)                      # allocated right behind da on most allocators; checked loosely
    ops.conv_dgrad_tf32(dy, wd, da, B=B, Tin=Tin, Tout=Tout, Cin=cin, N=cout, k=k, stride=s)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    ref = torch.nn.functional.conv_transpose1d(dy.double().transpose(1, 2), w.double(), stride=s)          # [B, cin, (Tout-1)*s + k]
    full = torch.zeros(B, cin, Tin, dtype=torch.float64, device=dev)
    full[:, :, :ref.shape[2]] = ref
    assert torch.isfinite(da).all()
    assert rel_l2(da, full.transpose(1, 2)) < TF32_TOL, rel_l2(da, full.transpose(1, 2))
    assert (guard == 7.0).all()


@pytest.mark.parametrize('M', [4352, 13056, 77])
def test_gemm_tf32_padded_pitch_head(dev, M):
    """The output head's 150-wide hidden layer with a 152-float row pitch: N % 4 != 0 with a vector epilogue (the last quad of a row is
    stored as scalars, the pad columns are never written), then K = 150 read through lda = ldb = 152 (the TMA maps end at column 150)."""
    from tgb200 import ops
    H, Hh, ld, D = 300, 150, 152, 27
    x = _rand(M, H, dev=dev); w0 = _rand(Hh, H, dev=dev, seed=1, scale=H ** -0.5); b0 = _rand(Hh, dev=dev, seed=2)
    y1 = torch.full((M, ld), 7.0, device=dev)
    ops.gemm_tf32(x, w0, y1, M=M, N=Hh, K=H, ldc=ld, bias=b0)
    ref1 = x.double() @ w0.double().t() + b0.double()
    assert rel_l2(y1[:, :Hh], ref1) < TF32_TOL, rel_l2(y1[:, :Hh], ref1)
    assert (y1[:, Hh:] == 7.0).all()
    w2p = torch.full((D, ld), float('nan'), device=dev); w2 = _rand(D, Hh, dev=dev, seed=3, scale=Hh ** -0.5); w2p[:, :Hh] = w2
    b2 = _rand(D, dev=dev, seed=4)
    y1[:, Hh:] = float('nan')                 # pad columns must not be read
    poses = torch.full((M, D), float('nan'), device=dev)
    ops.gemm_tf32(y1, w2p, poses, M=M, N=D, K=Hh, lda=ld, ldb=ld, bias=b2)
    ref2 = y1[:, :Hh].double() @ w2.double().t() + b2.double()
    assert rel_l2(poses, ref2) < TF32_TOL, rel_l2(poses, ref2)
    # data gradient through out.0: dhs = dy1 @ W0 with the padded transpose as the B operand
    w0tp = torch.full((2 * H, ld), float('nan'), device=dev); w0tp[:H, :Hh] = w0.t(); w0tp[H:, :Hh] = w0.t()      # stacked twice, as the engine does
    dout = torch.full((M, 2 * H), float('nan'), device=dev)
    ops.gemm_tf32(y1, w0tp, dout, M=M, N=2 * H, K=Hh, lda=ld, ldb=ld)
    ref3 = y1[:, :Hh].double() @ w0.double()
    assert rel_l2(dout[:, :H], ref3) < TF32_TOL and rel_l2(dout[:, H:], ref3) < TF32_TOL, (rel_l2(dout[:, :H], ref3), rel_l2(dout[:, H:], ref3))
