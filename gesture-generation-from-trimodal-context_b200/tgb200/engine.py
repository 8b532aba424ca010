"""Launch plans for the hot path: PoseGenerator and ConvDiscriminator forward + hand-derived backward, expressed as
sequences of C-ABI kernel launches (tgb200.ops) over pre-allocated workspaces.

Layout: every activation is channels-last ([B, T, C] == row-major [B*T, C]); the reference's [B, C, T] conv layout,
its chomp copy (tcn.py:13), its torch.cat / repeat of the GRU input (multimodal_context_net.py:139-153) and its
per-parameter gradient tensors do not exist here.  Parameters live in a flat arena (tgb200.arena).

Reference anchors: scripts/model/multimodal_context_net.py:9-28 (WavEncoder), :31-61 + scripts/model/tcn.py
(TextEncoderTCN), :110-160 (PoseGenerator.forward), :207-252 (ConvDiscriminator)."""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import torch

from . import _lib, config, ops
from .arena import ParamArena


class _Overlap:
    """Fork / join helper over a few auxiliary CUDA streams (per device).  `with side.on(i):` makes stream i wait for
    everything queued so far on the current stream and runs the body there; `side.join(i)` makes the current stream wait
    for stream i.  Disabled (body runs inline) without CUDA, in trace mode, or with TGB200_OVERLAP=0."""

    def __init__(self):
        self._streams = {}
        self._active = set()          # side streams with work forked since their last join

    def enabled(self):
        return config.overlap() and not _lib.TRACE_ONLY and torch.cuda.is_available()

    def _stream(self, device, i):
        key = (device.index, i)
        if key not in self._streams:
            # the per-block preparation streams (dropout masks, weight-normed filters of one TextEncoderTCN block) carry a few ~10 us
            # launches each that gate a block of the longest chain in front of the GRU: at default priority they were starved for 100+ us
            # by the wide kernels of the audio / discriminator branches (profiles/r02_timeline_step_final.txt, 57-510 us)
            # Priorities (kernel nodes of a captured graph inherit them): the text chain's head - embedding / block-0 masks and filters - goes
            # first, then the iteration's main chain (captured at -2, train_gan._capture_stream), then the later blocks' preparation; the
            # audio branch, weight gradients and the discriminator's real pass keep the default.  The top of the iteration is ~150 us of
            # bandwidth-bound kernels that each fill the GPU, so what starts first decides when the 8-GEMM text chain can start.
            prio = -3 if i in (S_PREP[0], S_PREP0W, S_WAVB) else (-1 if i in S_PREP else 0)
            if config.flat_prio() and prio < 0:
                prio = -1
            hi = min(torch.cuda.Stream.priority_range())         # numerically lowest = most urgent (torch exposes 0 .. -3)
            self._streams[key] = torch.cuda.Stream(device=device, priority=max(prio, hi))
        return self._streams[key]

    @contextlib.contextmanager
    def on(self, i):
        if not self.enabled():
            yield
            return
        cur = torch.cuda.current_stream()
        s = self._stream(cur.device, i)
        ev = torch.cuda.Event()
        ev.record(cur)
        s.wait_event(ev)
        self._active.add((cur.device.index, i))
        with torch.cuda.stream(s):
            yield

    def join(self, i):
        if not self.enabled():
            return
        if i == S_WGRAD:
            self.join(S_BIAS)             # the bias-gradient column sums forked off the weight-gradient stream
            self.join(S_BIAS2)
            self.join(S_WGRAD2)           # the second / third weight-gradient streams (independent filters round-robin over them)
            self.join(S_WGRAD3)
        cur = torch.cuda.current_stream()
        key = (cur.device.index, i)
        if key not in self._active or self._streams[key] == cur:
            return                        # nothing outstanding (also keeps un-forked streams out of a CUDA-graph capture)
        self._active.discard(key)
        ev = torch.cuda.Event()
        ev.record(self._streams[key])
        cur.wait_event(ev)


side = _Overlap()
S_WGRAD, S_WAV, S_DREAL, S_WAVW, S_SPK, S_BIAS, S_WGRAD2, S_WGRAD3, S_SCALARS = 1, 2, 3, 4, 5, 6, 7, 8, 9
S_WAVB = 17                    # WavEncoder backward: the longest chain behind the GRU backward, highest priority
S_ZERO = 16                    # zero fills of split-K outputs, ahead of their GEMMs
S_BIAS2 = 15                   # second bias-gradient stream: the column sums alternate between S_BIAS and S_BIAS2
S_PREP0W = 14                  # weight-normed filters of block 0 (beside its masks on S_PREP[0])
S_PREP = (10, 11, 12, 13)      # dropout masks + weight-normed filters of TextEncoderTCN block i live on S_PREP[i % 4], joined right before that block


def s_prep(i):
    return S_PREP[i % len(S_PREP)]

F32 = torch.float32
BN_EPS = 1e-5
BN_MOM = 0.1


class Workspace:
    """Named scratch tensors, allocated once per shape (so steady-state steps allocate nothing and can be graph-captured)."""

    def __init__(self, device):
        self.device = device
        self.t: Dict[str, torch.Tensor] = {}

    def get(self, name, shape, dtype=F32, zero=False):
        shape = tuple(int(s) for s in shape)
        t = self.t.get(name)
        if t is None or t.shape != shape or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, device=self.device, dtype=dtype)
            self.t[name] = t
        return t

    def __getitem__(self, name):
        return self.t[name]


def _tf32_ok(A, lda, W, ldw, K):
    """Operands a TMA descriptor can describe: 16-byte aligned base and pitch (K < 32 is zero-filled up to one 32-float swizzle atom
    by the TMA unit: a latency-bound small GEMM still beats the FFMA kernel's 128-row CTAs)."""
    return (config.fast() and K >= 8 and K % 4 == 0 and lda % 4 == 0 and ldw % 4 == 0 and A.data_ptr() % 16 == 0 and W.data_ptr() % 16 == 0)


def mm_nt(A, W, C, *, M, N, K, lda=None, **epi):
    """C[M,N] = epi(A[M,K] @ W[N,K]^T): tcgen05 TF32 GEMM in fast mode when the operands qualify, fp32 FFMA GEMM otherwise."""
    lda = K if lda is None else lda
    if _tf32_ok(A, lda, W, K, K):
        ops.gemm_tf32(A, W, C, M=M, N=N, K=K, lda=lda, **epi)
    else:
        ops.conv_gemm(A, W, C, B=1, Tin=M, Tout=M, N=N, Cin=K, taps=1, lda=lda, ldw=K, wsc=1, **epi)


def mm_nn(dY, W, Wt, dX, *, M, N, K, **epi):
    """dX[M,K] = epi(dY[M,N] @ W[N,K]).  Wt = W^T [K,N] (prepared in fast mode) feeds the tensor-core path."""
    if Wt is not None and _tf32_ok(dY, N, Wt, N, N):
        ops.gemm_tf32(dY, Wt, dX, M=M, N=K, K=N, **epi)
    else:
        ops.linear_dgrad(dY, W, dX, M=M, K=K, N=N, **epi)


_bias_rr = [0]


def wgrad(X, G, dW, *, B, T, N, Cin, shift=0, ldx=None, ldg=None, ldw=None, dbias=None):
    """dW[N,Cin] += sum_{b,t} G[(b,t), :]^T X[(b,t+shift), :] (rows outside the clip are zero): tcgen05 TF32 kernel in fast mode
    when TMA can describe the operands, fp32 FFMA split-K kernel otherwise."""
    ldx = Cin if ldx is None else ldx
    ldg = N if ldg is None else ldg
    if (config.fast() and Cin >= 8 and Cin % 4 == 0 and N >= 16 and ldx % 4 == 0 and ldg % 4 == 0 and X.data_ptr() % 16 == 0 and G.data_ptr() % 16 == 0
            and (shift == 0 or T <= 40)):
        if dbias is not None:
            # ~3 us column-sum launches: on their own stream they run beside the weight-gradient GEMMs instead of between them (60 per
            # iteration; they owned 89 us of the step when serialised on the weight-gradient stream, profiles/r02_timeline_step_ownership.txt)
            # two streams, alternating: the discriminator's twelve column sums in a row outlasted its weight-gradient GEMMs by ~25 us and
            # gated its Adam step (and, in the generator pass, the hand-over of d poses) - profiles/r02_timeline_step_3p86ms.txt, 1856-1920 us
            _bias_rr[0] ^= 1
            with side.on(S_BIAS2 if _bias_rr[0] else S_BIAS):
                ops.col_sum(G, ldg, B * T, N, dbias)
        ops.wgrad_tf32(G, X, dW, B=B, T=T, N=N, Cin=Cin, shift=shift, ldg=ldg, ldx=ldx, ldw=ldw, dbias=None)
    else:
        ops.conv_wgrad(X, G, dW, B=B, Tin=T, Tout=T, N=N, Cin=Cin, taps=1, pad=-shift, lda=ldx, ldg=ldg, ldw=Cin if ldw is None else ldw,
                       dbias=dbias)


def _conv_out(tin, k, stride, pad=0, dil=1):
    return (tin + 2 * pad - dil * (k - 1) - 1) // stride + 1


def gru_arena_order(names: List[str]) -> List[str]:
    """Arena order with weight_ih / bias_ih of both directions adjacent, so that one GEMM produces gi for fwd|rev."""
    gru = [n for n in names if n.startswith('gru.')]
    rest = [n for n in names if not n.startswith('gru.')]
    layers = sorted({int(n.split('_l')[1].split('_')[0]) for n in gru})
    out = []
    for l in layers:
        for kind in ('weight_ih', 'bias_ih', 'weight_hh', 'bias_hh'):
            out += [f'gru.{kind}_l{l}', f'gru.{kind}_l{l}_reverse']
    assert sorted(out) == sorted(gru)
    return rest + out


# =====================================================================================================================
# shared bidirectional multi-layer GRU plan
# =====================================================================================================================
class GruPlan:
    def __init__(self, arena: ParamArena, prefix: str, in_size: int, hidden: int, layers: int, ws: Workspace, tag: str):
        self.a, self.pre, self.I, self.H, self.L, self.ws, self.tag = arena, prefix, in_size, hidden, layers, ws, tag
        for l in range(layers):
            for kind in ('weight_ih', 'bias_ih'):
                assert arena.adjacent(f'{prefix}.{kind}_l{l}', f'{prefix}.{kind}_l{l}_reverse')

    def tc(self):
        """Tensor-core recurrence kernels are used in fast mode for the hidden sizes they support."""
        # H <= 64 (the discriminator) runs the single-CTA-per-tile fp32 kernels in both modes: no inter-CTA stepping at all
        return config.fast() and self.H % 4 == 0 and 64 < self.H <= 384

    def _w(self, kind, l, rev=False):
        return self.a.view(f'{self.pre}.{kind}_l{l}' + ('_reverse' if rev else ''))

    def _g(self, kind, l, rev=False):
        return self.a.gview(f'{self.pre}.{kind}_l{l}' + ('_reverse' if rev else ''))

    def prep(self):
        """W_hh^T for the forward recurrence kernel (weights change every optimiser step)."""
        H = self.H
        for l in range(self.L):
            for d in (0, 1):
                wt = self.ws.get(f'{self.tag}.whhT{l}_{d}', (H, 3 * H))
                ops.transpose(self._w('weight_hh', l, bool(d)), wt, 3 * H, H)
            K = self.I if l == 0 else 2 * H
            if config.fast() and K % 4 == 0 and K >= 8:
                # [6H, K] (both directions, adjacent in the arena) -> [K, 6H]: operand of the data-gradient GEMM
                ops.transpose(self._w('weight_ih', l), self.ws.get(f'{self.tag}.wihT{l}', (K, 6 * H)), 6 * H, K)

    def forward(self, x, B, T, masks: Optional[List[Optional[torch.Tensor]]], save: bool, hook=None, hook_after: int = 0, before_rec0=None):
        """x [B*T, I] -> out of the last layer [B*T, 2H].  masks[l] multiplies the output of layer l (l < L-1).
        hook() is called once the recurrence of layer `hook_after` has been queued: work forked there starts when that recurrence has
        finished, beside the next input projection.  before_rec0() is called between layer 0's projection and its recurrence: work forked
        there runs beside the first recurrence."""
        H, ws, tag = self.H, self.ws, self.tag
        M = B * T
        gi = ws.get(f'{tag}.gi', (M, 6 * H))
        tc = self.tc()
        sync = ws.get(f'{tag}.sync', (max(ops.gru_tf32_sync_ints(B, H) if tc else ops.gru_sync_ints(B, H), 1),), torch.int32)
        inp, K = x, self.I
        for l in range(self.L):
            mm_nt(inp, self._w('weight_ih', l), gi, M=M, N=6 * H, K=K, bias=self._w('bias_ih', l))
            out = ws.get(f'{tag}.out{l}', (M, 2 * H))
            saved = ws.get(f'{tag}.saved{l}', (4, M, 2 * H)) if save else None
            mk = masks[l] if (masks is not None and l < self.L - 1) else None
            drop = ws.get(f'{tag}.drop{l}', (M, 2 * H)) if mk is not None else None
            if l == 0 and before_rec0 is not None:
                before_rec0()
            if hook is not None and hook_after >= 100 and l == min(hook_after - 100, self.L - 1):
                hook()                      # 'preN': forked between layer N's projection and its recurrence (runnable together with it)
            if mk is not None and l > 0:
                side.join(s_prep(l + 1))    # this layer's mask, drawn beside the first recurrence (PoseGenerator engine: _late_masks)
            if tc and mk is not None:
                # the inter-layer dropout rides on the recurrence kernel's output store (was a separate 94 MB elementwise pass per layer)
                ops.gru_layer_fwd_tf32_drop(gi, self._w('weight_hh', l), self._w('weight_hh', l, True), self._w('bias_hh', l),
                                            self._w('bias_hh', l, True), out, saved, M * 2 * H, mk, drop, sync, B, T, H)
            elif tc:
                ops.gru_layer_fwd_tf32(gi, self._w('weight_hh', l), self._w('weight_hh', l, True), self._w('bias_hh', l),
                                       self._w('bias_hh', l, True), out, saved, M * 2 * H, sync, B, T, H)
            else:
                ops.gru_layer_fwd(gi, ws[f'{tag}.whhT{l}_0'], ws[f'{tag}.whhT{l}_1'], self._w('bias_hh', l), self._w('bias_hh', l, True),
                                  out, saved, M * 2 * H, sync, B, T, H)
            if hook is not None and hook_after < 100 and l == min(hook_after, self.L - 1):
                hook()
            if mk is not None:
                if not tc:
                    ops.mul(out, mk, drop, M * 2 * H)
                inp = drop
            else:
                inp = out
            K = 2 * H
        self._fwd_inputs = x
        return inp

    def weight_grads(self, l, inp, dgi, dgh, out, Bb, T, stream=S_WGRAD):
        """Weight / bias gradients of layer l from its gate gradients, on a weight-gradient stream (off the recurrence chain)."""
        H = self.H
        K = self.I if l == 0 else 2 * H
        with side.on(stream):
            wgrad(inp, dgi, self._g('weight_ih', l), B=Bb, T=T, N=6 * H, Cin=K, dbias=self._g('bias_ih', l))
            for d in (0, 1):
                # dW_hh[d] += dgh_d^T h_prev ; h_prev(t) = out(t-1) (fwd) / out(t+1) (rev), zero outside the clip
                wgrad(out[:, d * H:], dgh[:, d * 3 * H:], self._g('weight_hh', l, bool(d)), B=Bb, T=T, N=3 * H, Cin=H,
                      shift=(-1 if d == 0 else 1), ldx=2 * H, ldg=6 * H, dbias=self._g('bias_hh', l, bool(d)))

    def backward(self, dout, x, B_all, lo, hi, T, masks, need_dx: bool, join_first: bool = True):
        """dout [(hi-lo)*T, 2H] = gradient of the last layer's output for clips [lo,hi) of a forward over B_all clips.
        Accumulates all weight grads; returns d x [(hi-lo)*T, I] (or None)."""
        H, ws, tag = self.H, self.ws, self.tag
        Bb = hi - lo
        Mb, M_all = Bb * T, B_all * T
        r0, r1 = lo * T, hi * T
        tc = self.tc()
        if join_first:
            side.join(S_WGRAD)  # weight-gradient launches of an earlier backward may still be reading dgi / dgh
        partial = ws.get(f'{tag}.partial', (max(ops.gru_bwd_tf32_scratch_floats(Bb, H) if tc else ops.gru_bwd_scratch_floats(Bb, H), 1),))
        sync = ws.get(f'{tag}.bsync', (max(ops.gru_tf32_sync_ints(Bb, H) if tc else ops.gru_sync_ints(Bb, H), 1),), torch.int32)
        dx = None
        # The data-gradient GEMMs through W_ih are long-K, few-tile problems that run split-K with red.global.add; their outputs are zeroed
        # HERE, on a side stream, and the GEMMs accumulate: a memset node between the recurrence and its GEMM cost ~25 us of gaps per layer
        # on the serial chain (profiles/r02_timeline_step_3p78ms.txt, 2449-2479 us).  One buffer per layer, so all can be zeroed up front.
        with side.on(S_ZERO):
            for l in range(self.L - 1, -1, -1):
                if l > 0 or need_dx:
                    ws.get(f'{tag}.dx{l}' if l > 0 else f'{tag}.dxin', (Mb, self.I if l == 0 else 2 * H)).zero_()
        zero_joined = False
        for l in range(self.L - 1, -1, -1):
            dgi = ws.get(f'{tag}.dgi{l}', (Mb, 6 * H))      # per layer: the weight gradients consume them on a side stream
            dgh = ws.get(f'{tag}.dgh{l}', (Mb, 6 * H))
            out = ws[f'{tag}.out{l}'][r0:r1]
            saved = ws[f'{tag}.saved{l}'][:, r0:r1]          # plane stride stays M_all*2H
            if tc:
                ops.gru_layer_bwd_tf32(dout, out, saved[0], M_all * 2 * H, ws[f'{tag}.whhT{l}_0'], ws[f'{tag}.whhT{l}_1'], dgi, dgh,
                                       partial, sync, Bb, T, H)
            else:
                ops.gru_layer_bwd(dout, out, saved[0], M_all * 2 * H, self._w('weight_hh', l), self._w('weight_hh', l, True), dgi, dgh,
                                  partial, sync, Bb, T, H)
            K = self.I if l == 0 else 2 * H
            if l == 0:
                inp = x[r0:r1]
            elif masks is not None and masks[l - 1] is not None:
                inp = ws[f'{tag}.drop{l - 1}'][r0:r1]
            else:
                inp = ws[f'{tag}.out{l - 1}'][r0:r1]
            # The weight gradients of this layer are forked AFTER its data-gradient GEMM: queued before it, their 200+ CTAs flooded the SMs
            # the moment the GEMM drained and the next recurrence's 8-CTA clusters waited ~20 us for room; forked here they become runnable
            # together with the next recurrence, whose higher stream priority places its clusters first.
            late_w = config.wgrad_after_dgrad() and (l > 0 or need_dx)
            if not late_w:
                self.weight_grads(l, inp, dgi, dgh, out, Bb, T)
            if l > 0 or need_dx:
                dx = ws.get(f'{tag}.dx{l}' if l > 0 else f'{tag}.dxin', (Mb, K))
                m = masks[l - 1][r0:r1] if (l > 0 and masks is not None and masks[l - 1] is not None) else None
                wt = ws.t.get(f'{tag}.wihT{l}') if config.fast() else None
                if not zero_joined:
                    side.join(S_ZERO)
                    zero_joined = True
                mm_nn(dgi, self._w('weight_ih', l), wt, dx, M=Mb, N=6 * H, K=K, mask=m, accumulate=True)
                if late_w:
                    self.weight_grads(l, inp, dgi, dgh, out, Bb, T)
                dout = dx
            else:
                dx = None
        return dx


# =====================================================================================================================
# PoseGenerator
# =====================================================================================================================
class GeneratorEngine:
    WAV = ((1, 16, 15, 5, 1600), (16, 32, 15, 6, 0), (32, 64, 15, 6, 0), (64, 32, 15, 6, 0))   # Cin, Cout, k, stride, pad

    def __init__(self, module):
        self.m = module
        names = [n for n, _ in module.named_parameters()]
        self.arena = ParamArena(module, gru_arena_order(names))
        self.ws: Optional[Workspace] = None
        self.gru: Optional[GruPlan] = None
        self.slots = {}
        self.use_audio = module.input_context in ('both', 'audio')
        self.use_text = module.input_context in ('both', 'text')
        self.z_mode = module.z_mode            # 'speaker' | 'random' | None
        self.H = module.hidden_size
        self.L = module.gru.num_layers
        self.I = module.in_size
        self.E = module.text_encoder.embedding.weight.shape[1]
        self.n_tcn = len(module.text_encoder.tcn.network)
        self.tcn_k = module.text_encoder.tcn.network[0].conv1.weight_v.shape[2]
        self.p_emb = float(module.text_encoder.emb_dropout)
        self.p_tcn = float(module.text_encoder.tcn.network[0].dropout1.p)
        self.p_gru = float(module.gru.dropout)
        self.bufs = dict(module.named_buffers())

    # -------------------------------------------------------------------------------------------- setup
    def ensure(self, device, slot='default'):
        """Binds the flat arena and selects the workspace `slot` (distinct slots never share scratch memory, so a
        CUDA graph captured on one slot stays valid while another slot is used with other shapes)."""
        if not self.arena.is_current():
            self.slots = {}
        self.arena.ensure(device)
        key = (slot, str(device))
        if key not in self.slots:
            w = Workspace(device)
            self.slots[key] = (w, GruPlan(self.arena, 'gru', self.I, self.H, self.L, w, 'g'))
        self.ws, self.gru = self.slots[key]
        self.bufs = dict(self.m.named_buffers())
        return self

    def P(self, name):
        return self.arena.params[name].data

    def G(self, name):
        return self.arena.gview(name)

    def prep_weights(self, part: str = 'all'):
        """Per-optimiser-step derived weights: weight-norm'ed TCN filters ('tcn') and transposed recurrent / head matrices ('rest')."""
        ws = self.ws
        if self.use_text and part in ('all', 'tcn'):
            def norm_block(i):
                for j in (1, 2):
                    q = f'text_encoder.tcn.network.{i}.conv{j}'
                    v = self.P(q + '.weight_v')
                    N, Cin, k = v.shape
                    wT = ws.get(f'tcn.wT{i}_{j}', (k, Cin, N)) if config.fast() else None
                    ops.weight_norm_fwd(v, self.P(q + '.weight_g'), ws.get(f'tcn.w{i}_{j}', (k, N, Cin)), wT, ws.get(f'tcn.inv{i}_{j}', (N,)),
                                        N, Cin, k)
            with side.on(S_PREP0W):             # block 0: beside its masks (S_PREP[0]); text_forward joins both in front of block 0
                norm_block(0)
            for i in range(1, self.n_tcn):      # filters of the later blocks: each joined in text_forward right before its block
                with side.on(s_prep(i)):
                    norm_block(i)
        if part == 'tcn':
            return
        with side.on(S_WAV):        # not needed before the GRU / the backward pass: off the critical path (joined before the GRU input)
            if config.fast():
                for name in ('text_encoder.decoder.weight', 'out.0.weight', 'out.2.weight'):
                    if name.startswith('text') and not self.use_text:
                        continue
                    w = self.P(name)
                    ops.transpose(w, ws.get('T.' + name, (w.shape[1], w.shape[0])), w.shape[0], w.shape[1])
                if self.head_ld() != self.H // 2:
                    # the head's 150-wide hidden layer with a 152-float row pitch (TMA pitches are multiples of 16 bytes): padded copies of
                    # out.2.weight [D,150] and out.0.weight^T [H,150]; the pad columns are never read (the TMA maps end at column 150)
                    Hh, ld = self.H // 2, self.head_ld()
                    ws.get('P.out.2.weight', (self.m.pose_dim, ld), zero=True)[:, :Hh].copy_(self.P('out.2.weight'))
                    pt = ws.get('PT.out.0.weight', (2 * self.H, ld), zero=True)      # stacked twice: the data gradient lands in both
                    pt[:self.H, :Hh].copy_(self.P('out.0.weight').t())             # direction halves of d out in one GEMM
                    pt[self.H:, :Hh].copy_(self.P('out.0.weight').t())
            self.gru.prep()

    def head_ld(self):
        """Row pitch of the output head's hidden activation y1 [M, H/2]: padded to a multiple of 4 floats in fast mode so that both head
        GEMMs and the data gradient through out.0 run on the tensor cores (H/2 = 150: a 600-byte pitch is not a TMA pitch)."""
        Hh = self.H // 2
        return (Hh + 3) // 4 * 4 if (config.fast() and config.head_padded() and Hh >= 32) else Hh

    def make_masks(self, Bt, T, seed, offset_dev, sid0=0, split=False):
        """Dropout keep-masks (scaled by 1/(1-p)) for one training forward over Bt clips, from the Philox kernel.
        split=True: block i's masks (and the embedding's, with block 0) are drawn on that block's preparation stream
        (joined in text_forward right before block i), the GRU's on the weight-gradient stream (idle during the forward
        pass, joined in front of the GRU); split=False: everything on the current stream."""
        ws, M = self.ws, Bt * T
        masks = {}
        jobs = []                                   # (side stream or None, buffer, numel, p, Philox stream id)
        sid = sid0
        if self.use_text:
            if self.p_emb > 0:
                masks['emb'] = ws.get('mask.emb', (M, self.E)); jobs.append((s_prep(0), masks['emb'], M * self.E, self.p_emb, sid))
            sid += 1
            for i in range(self.n_tcn):
                for j in (1, 2):
                    if self.p_tcn > 0:
                        mk = ws.get(f'mask.tcn{i}_{j}', (M, self.H))
                        jobs.append((s_prep(i), mk, M * self.H, self.p_tcn, sid))
                        masks[f'tcn{i}_{j}'] = mk
                    sid += 1
        for l in range(self.L - 1):
            if self.p_gru > 0:
                mk = ws.get(f'mask.gru{l}', (M, 2 * self.H))
                jobs.append((S_WGRAD, mk, M * 2 * self.H, self.p_gru, sid))
                masks[f'gru{l}'] = mk
            sid += 1
        for st, mk, n, p, sd in jobs:
            if st is None or not split:
                ops.philox_dropout_mask(mk, n, p, seed, offset_dev, sd)
        self._late_masks = None
        if split:
            for stream in sorted({st for st, *_ in jobs if st is not None and st != S_WGRAD}):
                with side.on(stream):
                    for st, mk, n, p, sd in jobs:
                        if st == stream:
                            ops.philox_dropout_mask(mk, n, p, seed, offset_dev, sd)
            # the GRU's inter-layer masks (3 x 31 MB at batch 3 x 128): layer 0's here, at low urgency; the others are drawn by forward() on
            # the SMs the first recurrence leaves idle instead of at the bandwidth-saturated top of the iteration
            late = [(mk, n, p, sd) for st, mk, n, p, sd in jobs if st == S_WGRAD]
            if late:
                with side.on(s_prep(self.n_tcn - 1) if self.use_text else S_WGRAD):     # layer 0's mask: with the last TCN block's preparation
                    mk, n, p, sd = late[0]
                    ops.philox_dropout_mask(mk, n, p, seed, offset_dev, sd)
                self._late_masks = (late[1:], seed, offset_dev)
        return masks

    def start_wav(self, in_audio, training, n_bn_updates=1):
        """Queues the WavEncoder forward on its side stream ahead of everything else of the iteration (it is the longest
        chain in front of the GRU); forward() picks the result up instead of launching it again."""
        with side.on(S_WAV):
            self._wav_feat = self.wav_forward(in_audio, training, n_bn_updates)

    # -------------------------------------------------------------------------------------------- WavEncoder
    def wav_fast(self):
        """conv2-4 on the tensor cores (fast mode): every pad is 0 and every pitch is a multiple of 4 floats."""
        return config.fast() and all(p == 0 and c % 4 == 0 for (c, _, _, _, p) in self.WAV[1:])

    def wav_forward(self, audio, training: bool, n_updates: int = 1):
        """multimodal_context_net.py:9-28.  audio [Ba, L] -> [Ba*T, 32].  BatchNorm+LeakyReLU(0.3) are never materialised:
        they are applied as the next convolution's operand prologue."""
        ws = self.ws
        Ba, L = audio.shape
        pre = 'audio_encoder.feat_extractor.'
        self.wav_T = [L]
        x = audio
        scale = shift = None
        fast = self.wav_fast()
        for li, (cin, cout, k, s, pad) in enumerate(self.WAV):
            conv = pre + str(3 * li)
            tin = self.wav_T[-1]
            tout = _conv_out(tin, k, s, pad)
            self.wav_T.append(tout)
            y = ws.get(f'wav.y{li}', (Ba * tout, cout))
            if li == 0:
                ops.conv1_direct(x, self.P(conv + '.weight'), self.P(conv + '.bias'), y, B=Ba, Tin=tin, Tout=tout, N=cout, taps=k, stride=s, pad=pad)
            elif fast:
                # tensor cores: materialise lrelu(bn(y_prev)) once, then a TF32 GEMM whose A rows are the overlapping
                # windows [k*cin] of that channels-last activation (row pitch s*cin floats), read in place by TMA
                a = ws.get(f'wav.a{li - 1}', (Ba * tin, cin))
                ops.affine_lrelu(x, a, Ba * tin, cin, scale, shift, 0.3)
                w2 = ws.get(f'wav.w2_{li}', (cout, k * cin)); w2t = ws.get(f'wav.w2t_{li}', (k * cin, cout))
                ops.window_weights(self.P(conv + '.weight'), w2, w2t, cout, cin, k)
                if config.wav_dgrad_direct() and cout >= 8:
                    wd = ws.get(f'wav.wd_{li}', (-(-k // s) * s * cin, cout))
                    with side.on(S_WAVW):           # only the backward reads it (wav_backward runs after the forward's joins)
                        ops.window_dgrad_weights(self.P(conv + '.weight'), wd, cout, cin, k, s)
                ops.gemm_tf32(a, w2, y, M=Ba * tout, N=cout, K=k * cin, lda=s * cin, clip_rows=tout, a_clip_pitch=tin * cin,
                              bias=self.P(conv + '.bias'))
            else:
                ops.conv1d(x, self.P(conv + '.weight'), self.P(conv + '.bias'), y, B=Ba, Tin=tin, Cin=cin, N=cout, k=k, stride=s, pad=pad,
                           pscale=scale, pshift=shift, pslope=0.3)
            if li < 3:
                bn = pre + str(3 * li + 1)
                mean, rstd = ws.get(f'wav.mean{li}', (cout,)), ws.get(f'wav.rstd{li}', (cout,))
                scale, shift = ws.get(f'wav.scale{li}', (cout,)), ws.get(f'wav.shift{li}', (cout,))
                if training:
                    sums = ws.get(f'wav.sums{li}', (2 * cout,), torch.float64)
                    sums.zero_()
                    ops.col_stats(y, cout, Ba * tout, cout, sums)
                    ops.bn_finalize(sums, Ba * tout, cout, BN_EPS, BN_MOM, n_updates, self.P(bn + '.weight'), self.P(bn + '.bias'),
                                    self.bufs[bn + '.running_mean'], self.bufs[bn + '.running_var'], self.bufs[bn + '.num_batches_tracked'],
                                    mean, rstd, scale, shift)
                else:
                    ops.bn_eval_fold(self.P(bn + '.weight'), self.P(bn + '.bias'), self.bufs[bn + '.running_mean'],
                                     self.bufs[bn + '.running_var'], BN_EPS, None, scale, shift, cout)
            x = y
        return x                                           # [Ba*34, 32]

    def wav_backward(self, d_feat, audio):
        """d_feat [Ba*T, 32] -> accumulates conv / BN grads (train-mode BN)."""
        ws = self.ws
        Ba = audio.shape[0]
        pre = 'audio_encoder.feat_extractor.'
        dy = d_feat
        fast = self.wav_fast()
        for li in (3, 2, 1, 0):
            cin, cout, k, s, pad = self.WAV[li]
            conv = pre + str(3 * li)
            tin, tout = self.wav_T[li], self.wav_T[li + 1]
            x = audio if li == 0 else ws[f'wav.y{li - 1}']
            sc = ws[f'wav.scale{li - 1}'] if li > 0 else None
            sh = ws[f'wav.shift{li - 1}'] if li > 0 else None
            if fast and li > 0:
                a = ws[f'wav.a{li - 1}']
                with side.on(S_WAVW):               # weight gradient off the data-gradient chain
                    dw2 = ws.get(f'wav.dw2_{li}', (cout, k * cin)); dw2.zero_()
                    ops.wgrad_tf32(dy, a, dw2, B=Ba, T=tout, N=cout, Cin=k * cin, ldx=s * cin, x_clip_pitch=tin * cin,
                                   dbias=self.G(conv + '.bias'))
                    ops.window_wgrad_add(dw2, self.G(conv + '.weight'), cout, cin, k)
                da = ws.get(f'wav.da{li - 1}', (Ba * tin, cin))
                if config.wav_dgrad_direct() and cout >= 8:
                    # transposed convolution as ceil(k/s) accumulating GEMM taps over dy rows shifted in place by TMA: no column matrix
                    # (the column GEMM + col2im wrote and re-read 161 MB for conv2: 188 us of the backward's critical tail)
                    ops.conv_dgrad_tf32(dy, ws[f'wav.wd_{li}'], da, B=Ba, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k, stride=s)
                else:
                    col = ws.get('wav.col', (Ba * self.wav_T[2] * self.WAV[1][2] * self.WAV[1][0],))[:Ba * tout * k * cin].view(Ba * tout, k * cin)
                    ops.gemm_tf32(dy, ws[f'wav.w2t_{li}'], col, M=Ba * tout, N=k * cin, K=cout)
                    ops.col2im(col, da, B=Ba, Tin=tin, Tout=tout, Cin=cin, k=k, stride=s)
            elif fast and cin == 1 and cout == 16 and k <= 15:
                ops.conv1_wgrad(x, dy, self.G(conv + '.weight'), self.G(conv + '.bias'), B=Ba, Tin=tin, Tout=tout, N=cout, taps=k, stride=s,
                                pad=pad)
                break
            else:
                ops.conv1d_wgrad(x, dy, self.G(conv + '.weight'), self.G(conv + '.bias'), B=Ba, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k,
                                 stride=s, pad=pad, pscale=sc, pshift=sh, pslope=0.3)
                if li == 0:
                    break
                da = ws.get(f'wav.da{li - 1}', (Ba * tin, cin))
                ops.conv1d_dgrad(dy, self.P(conv + '.weight'), da, B=Ba, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k, stride=s, pad=pad)
            bn = pre + str(3 * (li - 1) + 1)
            sums = ws.get(f'wav.bsums{li - 1}', (2 * cin,), torch.float64)
            sums.zero_()
            mean, rstd = ws[f'wav.mean{li - 1}'], ws[f'wav.rstd{li - 1}']
            ops.bn_bwd_reduce(da, x, Ba * tin, cin, mean, rstd, sc, sh, 0.3, sums)
            ops.bn_bwd_apply(da, x, da, Ba * tin, cin, mean, rstd, sc, sh, 0.3, self.P(bn + '.weight'), sums, self.G(bn + '.weight'),
                             self.G(bn + '.bias'))
            dy = da
        side.join(S_WAVW)

    # -------------------------------------------------------------------------------------------- TextEncoderTCN
    def text_forward(self, in_text, Bt, T, masks):
        """multimodal_context_net.py:57-61 + tcn.py:43-64.  in_text [Ba, T] int64 (read at row % (Ba*T)) -> [Bt*T, 32]."""
        ws, H, E = self.ws, self.H, self.E
        M = Bt * T
        Ba = in_text.shape[0]
        idx_mod = Ba * T if Bt != Ba else 0
        emb = ws.get('txt.emb', (M, E))
        side.join(s_prep(0))                       # the embedding's and block 0's dropout masks (make_masks split=True)
        side.join(S_PREP0W)                        # block 0's weight-normed filters (prep_weights)
        ops.embedding_gather(self.P('text_encoder.embedding.weight'), in_text, idx_mod, masks.get('emb') if masks else None, emb, M, E)
        x, cin = emb, E
        k = self.tcn_k
        fused = self._tcn_fused = bool(config.fast() and config.tcn_fused_add())
        for i in range(self.n_tcn):
            if i >= 1:
                side.join(s_prep(i))               # this block's dropout masks (make_masks split=True) and weight-normed filters (prep_weights)
            d = 2 ** i
            q = f'text_encoder.tcn.network.{i}'
            assert cin == H, 'TemporalBlock.downsample (n_inputs != n_outputs) is not on the configured path (tcn.py:33)'
            y1 = ws.get(f'txt.y1_{i}', (M, H)); y2 = ws.get(f'txt.y2_{i}', (M, H)); xo = ws.get(f'txt.x{i}', (M, H))
            m1 = masks.get(f'tcn{i}_1') if masks else None
            m2 = masks.get(f'tcn{i}_2') if masks else None
            self._tcn_conv(x, ws[f'tcn.w{i}_1'], self.P(q + '.conv1.bias'), y1, Bt, T, cin, H, k, d, act1=ops.ACT_RELU, mask=m1)
            if fused:
                # xo = relu(relu(conv2 + b) * m2 + x) in conv2's epilogue: no y2 round trip, no add kernel on the 8-GEMM chain
                self._tcn_conv(y1, ws[f'tcn.w{i}_2'], self.P(q + '.conv2.bias'), xo, Bt, T, H, H, k, d, act1=ops.ACT_RELU, mask=m2, residual=x,
                               act2=ops.ACT_RELU)
            else:
                self._tcn_conv(y1, ws[f'tcn.w{i}_2'], self.P(q + '.conv2.bias'), y2, Bt, T, H, H, k, d, act1=ops.ACT_RELU, mask=m2)
                ops.add(y2, x, xo, M * H, relu=True)
            x, cin = xo, H
        feat = ws.get('txt.feat', (M, 32))
        mm_nt(x, self.P('text_encoder.decoder.weight'), feat, M=M, N=32, K=H, bias=self.P('text_encoder.decoder.bias'))
        return feat

    @staticmethod
    def _tcn_conv(x, w_tap, bias, y, B, T, cin, cout, k, d, **epi):
        """Causal dilated conv (tcn.py:19-31) on tap-major weights w_tap [k][cout][cin]; x [B*T, cin] -> y [B*T, cout]."""
        if k == 2 and _tf32_ok(x, cin, w_tap, cin, cin):
            ops.gemm_tf32(x, w_tap, y, M=B * T, N=cout, K=cin, taps=2, shift0=-d, T=T, bias=bias, **epi)
        else:
            ops.conv_gemm(x, w_tap, y, B=B, Tin=T, Tout=T, N=cout, Cin=cin, taps=k, dil=d, pad=(k - 1) * d, ldw=cin, wsj=cout * cin, wsc=1,
                          bias=bias, **epi)

    @staticmethod
    def _tcn_wgrad(x, dy, dw_tap, dbias, B, T, cin, cout, k, d):
        """dw_tap[j][n][c] += sum dy[(b,t), n] x[(b, t - (k-1-j)*d), c] for every tap j (tap-major gradient buffer)."""
        for j in range(k):
            wgrad(x, dy, dw_tap[j * cout * cin:], B=B, T=T, N=cout, Cin=cin, shift=-(k - 1 - j) * d, dbias=dbias if j == 0 else None)

    @staticmethod
    def _tcn_dgrad(dy, w_tap, wT_tap, dx, B, T, cin, cout, k, d, **epi):
        """Data gradient of _tcn_conv: dx[t] = sum_j dy[t + (k-1-j)*d] W_j (anti-causal)."""
        if k == 2 and wT_tap is not None and _tf32_ok(dy, cout, wT_tap, cout, cout):
            ops.gemm_tf32(dy, wT_tap, dx, M=B * T, N=cin, K=cout, taps=2, shift0=d, T=T, **epi)
        else:
            ops.conv_gemm(dy, w_tap, dx, B=B, Tin=T, Tout=T, N=cin, Cin=cout, taps=k, dil=-d, pad=-(k - 1) * d, ldw=1, wsj=cout * cin,
                          wsc=cin, **epi)

    def text_backward(self, d_feat, in_text, lo, hi, T, masks):
        ws, H, E = self.ws, self.H, self.E
        Bb = hi - lo
        Mb = Bb * T
        r0, r1 = lo * T, hi * T
        k = self.tcn_k
        sl = lambda t: t[r0:r1] if t is not None else None
        xl = ws[f'txt.x{self.n_tcn - 1}'][r0:r1]
        wgrad(xl, d_feat, self.G('text_encoder.decoder.weight'), B=Bb, T=T, N=32, Cin=H, dbias=self.G('text_encoder.decoder.bias'))
        dx = ws.get('txt.dA', (Mb, H)); dpre = ws.get('txt.dB', (Mb, H)); dy1 = ws.get('txt.dD', (Mb, H))
        # (no join of the weight-gradient streams here: every buffer they read is per convolution, and in train_iter_gan this point sits
        # behind the GRU's weight gradients AND the early Adam step of the recurrent range - a join made the text chain wait ~100 us for them)
        mm_nn(d_feat, self.P('text_encoder.decoder.weight'), ws.t.get('T.text_encoder.decoder.weight') if config.fast() else None, dx,
              M=Mb, N=32, K=H)
        for i in range(self.n_tcn - 1, -1, -1):
            d = 2 ** i
            q = f'text_encoder.tcn.network.{i}'
            xin = (ws[f'txt.x{i - 1}'] if i > 0 else ws['txt.emb'])[r0:r1]
            cin = H if i > 0 else E
            y1, y2, xo = ws[f'txt.y1_{i}'][r0:r1], ws[f'txt.y2_{i}'][r0:r1], ws[f'txt.x{i}'][r0:r1]
            m1 = sl(masks.get(f'tcn{i}_1')) if masks else None
            m2 = sl(masks.get(f'tcn{i}_2')) if masks else None
            dc2 = ws.get(f'txt.dc2_{i}', (Mb, H)); dc1 = ws.get(f'txt.dc1_{i}', (Mb, H))   # per conv: read by the side stream
            if getattr(self, '_tcn_fused', False):
                ops.tcn_res_bwd(dx, xo, xin, m2, dpre, dc2, Mb * H)            # final ReLU, dropout2 + relu2 in one pass (y2 was not stored)
            else:
                ops.relu_mask_bwd(dx, xo, None, dpre, Mb * H)                  # through the block's final ReLU
                ops.relu_mask_bwd(dpre, y2, m2, dc2, Mb * H)                   # dropout2 + relu2
            wT = lambda j: ws.t.get(f'tcn.wT{i}_{j}') if config.fast() else None
            # each filter's gradient chain (zero, two tap GEMMs, weight-norm backward: ~50 us of small launches) on one of three streams:
            # serialised on a single stream the eight chains outlasted the data-gradient chain by ~120 us at the end of the iteration
            with side.on((S_WGRAD, S_WGRAD2, S_WGRAD3)[(2 * i) % 3]):
                dw = ws.get(f'tcn.dw{i}_2', (k * H * H,)); dw.zero_()                 # tap-major [k][H][cin]
                self._tcn_wgrad(y1, dc2, dw, self.G(q + '.conv2.bias'), Bb, T, H, H, k, d)
                ops.weight_norm_bwd(dw, self.P(q + '.conv2.weight_v'), self.P(q + '.conv2.weight_g'), ws[f'tcn.inv{i}_2'],
                                    self.G(q + '.conv2.weight_v'), self.G(q + '.conv2.weight_g'), H, H, k)
            self._tcn_dgrad(dc2, ws[f'tcn.w{i}_2'], wT(2), dy1, Bb, T, H, H, k, d)
            ops.relu_mask_bwd(dy1, y1, m1, dc1, Mb * H)                        # dropout1 + relu1
            with side.on((S_WGRAD, S_WGRAD2, S_WGRAD3)[(2 * i + 1) % 3]):
                dw = ws.get(f'tcn.dw{i}_1', (k * H * cin,)); dw.zero_()
                self._tcn_wgrad(xin, dc1, dw, self.G(q + '.conv1.bias'), Bb, T, cin, H, k, d)
                ops.weight_norm_bwd(dw, self.P(q + '.conv1.weight_v'), self.P(q + '.conv1.weight_g'), ws[f'tcn.inv{i}_1'],
                                    self.G(q + '.conv1.weight_v'), self.G(q + '.conv1.weight_g'), H, cin, k)
            # d x_in = conv1^T(dc) + residual branch (dpre)
            self._tcn_dgrad(dc1, ws[f'tcn.w{i}_1'], wT(1), dx, Bb, T, cin, H, k, d, residual=dpre)
        emb_p = self.arena.params['text_encoder.embedding.weight']
        if emb_p.requires_grad:
            Ba = in_text.shape[0]
            assert Bb == Ba and lo % Ba == 0
            ops.embedding_scatter_add(dx, in_text, sl(masks.get('emb')) if masks else None, self.G('text_encoder.embedding.weight'), Mb, E)
        side.join(S_WGRAD)      # the filters' gradient chains forked above

    # -------------------------------------------------------------------------------------------- full forward
    def forward(self, pre_seq, in_text, in_audio, vid, eps, Bt, training, masks=None, n_bn_updates=1, save=True):
        """PoseGenerator.forward (multimodal_context_net.py:110-160) over Bt clips whose pre_seq/text/audio repeat with
        period Ba = in_audio.shape[0] (train_iter_gan runs its three generator passes as ONE Bt = 3*Ba pass).
        Returns views (poses [Bt,T,D], z, mu, logvar) into the workspace."""
        ws, m = self.ws, self.m
        Ba, T = pre_seq.shape[0], pre_seq.shape[1]
        Dp = pre_seq.shape[2]
        M = Bt * T
        self.ctx = dict(Bt=Bt, Ba=Ba, T=T, masks=masks, pre_seq=pre_seq, in_text=in_text, in_audio=in_audio, vid=vid, eps=eps)
        audio_feat = None
        if self.use_audio:
            audio_feat = getattr(self, '_wav_feat', None)
            self._wav_feat = None
        wav_late = self.use_audio and audio_feat is None and self.use_text and not config.wav_first()
        if self.use_audio and audio_feat is None and not wav_late:
            with side.on(S_WAV):            # the audio encoder is independent of the text / speaker branches
                audio_feat = self.wav_forward(in_audio, training, n_bn_updates)
        z = mu = logvar = None
        Z = 0
        if self.z_mode == 'speaker':
            Z = 16
            e0, e1 = ws.get('spk.e0', (Bt, Z)), ws.get('spk.e1', (Bt, Z))
            mu, logvar, z = ws.get('spk.mu', (Bt, Z)), ws.get('spk.logvar', (Bt, Z)), ws.get('spk.z', (Bt, Z))
            with side.on(S_SPK):            # a chain of tiny launches: off the text-encoder chain
                ops.embedding_gather(self.P('speaker_embedding.0.weight'), vid, 0, None, e0, Bt, Z)
                ops.linear(e0, self.P('speaker_embedding.1.weight'), self.P('speaker_embedding.1.bias'), e1, M=Bt, K=Z, N=Z)
                ops.linear(e1, self.P('speaker_mu.weight'), self.P('speaker_mu.bias'), mu, M=Bt, K=Z, N=Z)
                ops.linear(e1, self.P('speaker_logvar.weight'), self.P('speaker_logvar.bias'), logvar, M=Bt, K=Z, N=Z)
                ops.reparam_fwd(mu, logvar, eps, z, Bt * Z)
        elif self.z_mode == 'random':
            Z = 16
            z = eps
        text_feat = self.text_forward(in_text, Bt, T, masks) if self.use_text else None
        if wav_late:
            # queued AFTER the text chain: the TextEncoderTCN chain (8 dependent GEMMs, ~0.5 ms) is the longer of the two in front of the
            # GRU, and a replayed CUDA graph feeds the GPU its nodes in capture order - whatever is captured first starts first
            with side.on(S_WAV):
                audio_feat = self.wav_forward(in_audio, training, n_bn_updates)
        side.join(S_SPK)
        side.join(S_WAV)
        side.join(S_WGRAD)                          # the GRU's masks (make_masks split=True)
        side.join(S_WAVW)
        for st in S_PREP + (S_PREP0W,):
            side.join(st)                           # no-ops after text_forward; only live when the text branch is configured off
        in_data = ws.get('g.in', (M, self.I))
        Da = 32 if self.use_audio else 0
        Dt = 32 if self.use_text else 0
        assert Dp + Da + Dt + Z == self.I
        ops.gru_input_concat(pre_seq, audio_feat, text_feat, z, in_data, Bt, Ba, T, Dp, Da, Dt, Z)
        gmasks = [masks.get(f'gru{l}') for l in range(self.L)] if masks else None
        # work that does not depend on the generator (train_iter_gan: the discriminator's pass over the real clips) can be forked here
        late = getattr(self, '_late_masks', None)
        self._late_masks = None

        def draw_late():
            # masks of layers >= 1, one stream each (the TCN blocks' preparation streams are idle by now): forked between layer 0's
            # projection and its recurrence, so they are drawn on the SMs the first recurrence leaves idle; layer l joins only its own
            for l, (mk, n, p, sd) in enumerate(late[0]):
                with side.on(s_prep(l + 2)):
                    ops.philox_dropout_mask(mk, n, p, late[1], late[2], sd)
        hook, at = getattr(self, 'beside_gru', None), getattr(self, 'beside_gru_at', -1)
        self.beside_gru = None
        if hook is not None and at < 0:
            hook()
            hook = None
        out = self.gru.forward(in_data, Bt, T, gmasks, save, hook=hook, hook_after=at, before_rec0=draw_late if (late and late[0]) else None)
        H = self.H
        ld = self.head_ld()
        hsum = ws.get('g.hsum', (M, H)); y1 = ws.get('g.y1', (M, ld), zero=True); poses = ws.get('g.poses', (M, m.pose_dim))
        ops.sum_halves(out, hsum, M, H)
        if ld != H // 2:
            ops.gemm_tf32(hsum, self.P('out.0.weight'), y1, M=M, N=H // 2, K=H, ldc=ld, bias=self.P('out.0.bias'))
            ops.gemm_tf32(y1, ws['P.out.2.weight'], poses, M=M, N=m.pose_dim, K=H // 2, lda=ld, ldb=ld, bias=self.P('out.2.bias'))
        else:
            mm_nt(hsum, self.P('out.0.weight'), y1, M=M, N=H // 2, K=H, bias=self.P('out.0.bias'))      # LeakyReLU(True) == identity
            mm_nt(y1, self.P('out.2.weight'), poses, M=M, N=m.pose_dim, K=H // 2, bias=self.P('out.2.bias'))
        return poses.view(Bt, T, m.pose_dim), z, mu, logvar

    def backward(self, d_poses, lo, hi, d_mu=None, d_logvar=None, d_z=None):
        """Gradient of clips [lo,hi) of the last forward.  d_poses [(hi-lo),T,D]; d_mu/d_logvar/d_z [(hi-lo),16] optional
        (d_mu / d_logvar are used as accumulators and overwritten).  Accumulates into the flat gradient arena."""
        ws, m, c = self.ws, self.m, self.ctx
        T, Ba, Bt, masks = c['T'], c['Ba'], c['Bt'], c['masks']
        Bb = hi - lo
        Mb = Bb * T
        r0, r1 = lo * T, hi * T
        H, D = self.H, m.pose_dim
        d_poses = d_poses.reshape(Mb, D)
        Hh, ld = H // 2, self.head_ld()
        dy1 = ws.get('g.dy1', (Mb, ld), zero=True); dhs = ws.get('g.dhs', (Mb, H)); dout = ws.get('g.dout', (Mb, 2 * H))
        side.join(S_WGRAD)      # weight-gradient launches of an earlier backward may still be reading the scratch buffers below
        ops.linear_dgrad(d_poses, self.P('out.2.weight'), dy1, M=Mb, K=Hh, N=D, ldc=ld)          # K = 27 reduction: fp32 kernel
        if ld != Hh:
            # d out = [d hsum, d hsum] with d hsum = dy1 @ out.0.weight, on the tensor cores: A = dy1 (152-float pitch, K = 150), B = the padded
            # transpose stacked twice [2H, 152] - the gradient of the sum over directions rides on the GEMM instead of a duplication pass
            ops.gemm_tf32(dy1, ws['PT.out.0.weight'], dout, M=Mb, N=2 * H, K=Hh, lda=ld, ldb=ld)
        else:
            mm_nn(dy1, self.P('out.0.weight'), ws.t.get('T.out.0.weight') if config.fast() else None, dhs, M=Mb, N=Hh, K=H)
            ops.dup_halves(dhs, dout, Mb, H)
        # The head's weight gradients are fp32 CUDA-core kernels (27 / 150 columns, ~600 CTAs): forked HERE, behind the data-gradient chain,
        # they become runnable together with the last layer's recurrence, whose more urgent stream places its clusters first; forked
        # before it they held the SMs when the recurrence was launched and it started ~24 us late.
        with side.on(S_WGRAD):
            ops.linear_wgrad(ws['g.y1'][r0:r1], d_poses, self.G('out.2.weight'), self.G('out.2.bias'), M=Mb, K=Hh, N=D, lda=ld)
        with side.on(S_WGRAD2):
            # with the 152-float pitch of dy1 this is a tensor-core weight gradient too (was a 585-CTA fp32 kernel of ~45 us whose CTAs
            # fragmented the GPCs exactly when the first recurrence needed 8 whole clusters)
            wgrad(ws['g.hsum'][r0:r1], dy1, self.G('out.0.weight'), B=Bb, T=T, N=Hh, Cin=H, ldg=ld, dbias=self.G('out.0.bias'))
        gmasks = [masks.get(f'gru{l}') for l in range(self.L)] if masks else None
        need_dx = self.use_audio or self.use_text or self.z_mode == 'speaker'
        d_in = self.gru.backward(dout, ws['g.in'], Bt, lo, hi, T, gmasks, need_dx, join_first=False)      # joined above, before the head's forks
        if getattr(self, 'on_gru_grads', None) is not None:
            # data parallel: the recurrent layers' gradients (the tail of the flat arena, 22 MB of 53) are complete once the weight-gradient
            # stream has drained what gru.backward queued on it - their all-reduce starts now, under the text / audio encoder backward
            with side.on(S_WGRAD):
                side.join(S_BIAS)       # the bias gradients' column sums run on their own streams
                side.join(S_BIAS2)
                self.on_gru_grads()
        if not need_dx:
            side.join(S_WGRAD)
            return
        Dp = c['pre_seq'].shape[2]
        Da = 32 if self.use_audio else 0
        Dt = 32 if self.use_text else 0
        Z = 16 if self.z_mode else 0
        d_audio = ws.get('g.daudio', (Mb, max(Da, 1))); d_text = ws.get('g.dtext', (Mb, max(Dt, 1))); dzc = ws.get('g.dz', (Bb, max(Z, 1)))
        ops.gru_input_split_bwd(d_in, d_audio, d_text, dzc, Bb, T, Dp, Da, Dt, Z)
        if self.z_mode == 'speaker':
            spk_ctx = side.on(S_SPK)        # ten ~10 us launches: on the main chain they held back the audio / text encoder backward by ~85 us
        else:
            spk_ctx = contextlib.nullcontext()
        with spk_ctx:
            if self.z_mode == 'speaker':
                Zs = 16
                if d_z is not None:
                    ops.add(dzc, d_z, dzc, Bb * Zs)
                dmu = d_mu if d_mu is not None else ws.get('spk.dmu', (Bb, Zs), zero=True)
                dlv = d_logvar if d_logvar is not None else ws.get('spk.dlv', (Bb, Zs), zero=True)
                if d_mu is None: dmu.zero_()
                if d_logvar is None: dlv.zero_()
                ops.reparam_bwd(dzc, ws['spk.logvar'][lo:hi], c['eps'][lo:hi], dmu, dlv, Bb * Zs)
                e0, e1 = ws['spk.e0'][lo:hi], ws['spk.e1'][lo:hi]
                de1, de0 = ws.get('spk.de1', (Bb, Zs)), ws.get('spk.de0', (Bb, Zs))
                ops.linear_wgrad(e1, dmu, self.G('speaker_mu.weight'), self.G('speaker_mu.bias'), M=Bb, K=Zs, N=Zs)
                ops.linear_wgrad(e1, dlv, self.G('speaker_logvar.weight'), self.G('speaker_logvar.bias'), M=Bb, K=Zs, N=Zs)
                ops.linear_dgrad(dmu, self.P('speaker_mu.weight'), de1, M=Bb, K=Zs, N=Zs)
                ops.linear_dgrad(dlv, self.P('speaker_logvar.weight'), de1, M=Bb, K=Zs, N=Zs, accumulate=True)
                ops.linear_wgrad(e0, de1, self.G('speaker_embedding.1.weight'), self.G('speaker_embedding.1.bias'), M=Bb, K=Zs, N=Zs)
                ops.linear_dgrad(de1, self.P('speaker_embedding.1.weight'), de0, M=Bb, K=Zs, N=Zs)
                ops.embedding_scatter_add(de0, c['vid'][lo:hi], None, self.G('speaker_embedding.0.weight'), Bb, Zs)
        if self.use_audio:
            assert Bb == Ba and lo % Ba == 0
            # ten dependent kernels (~340 us): longer than the text encoder's backward, so it gets the more urgent stream of the two
            with side.on(S_WAVB):
                self.wav_backward(d_audio, c['in_audio'])
        if self.use_text:
            self.text_backward(d_text, c['in_text'], lo, hi, T, masks)
        side.join(S_WAVB)
        side.join(S_SPK)
        side.join(S_WGRAD)


# =====================================================================================================================
# ConvDiscriminator
# =====================================================================================================================
class DiscriminatorEngine:
    CONVS = (('pre_conv.0', 'pre_conv.1'), ('pre_conv.3', 'pre_conv.4'), ('pre_conv.6', None))

    def __init__(self, module):
        self.m = module
        names = [n for n, _ in module.named_parameters()]
        self.arena = ParamArena(module, gru_arena_order(names))
        self.ws: Optional[Workspace] = None
        self.gru: Optional[GruPlan] = None
        self.slots = {}
        self.H = module.hidden_size
        self.L = module.gru.num_layers
        self.p_gru = float(module.gru.dropout)

    def ensure(self, device, slot='default'):
        if not self.arena.is_current():
            self.slots = {}
        self.arena.ensure(device)
        key = (slot, str(device))
        if key not in self.slots:
            w = Workspace(device)
            self.slots[key] = (w, GruPlan(self.arena, 'gru', 8, self.H, self.L, w, 'd'))
        self.ws, self.gru = self.slots[key]
        self.bufs = dict(self.m.named_buffers())
        return self

    def P(self, name):
        return self.arena.params[name].data

    def G(self, name):
        return self.arena.gview(name)

    def prep_weights(self):
        # the fused stack reads the arena's GRU block directly; the per-layer plan's transposed recurrent matrices (12 launches per
        # optimiser step) are produced on first use by a forward that does not take the fused path
        self._gru_prep_stale = True

    def make_masks(self, B, T, seed, offset_dev, sid0=0, tag=''):
        masks = {}
        M = B * T
        for l in range(self.L - 1):
            if self.p_gru > 0:
                mk = self.ws.get(f'mask{tag}.gru{l}', (M, 2 * self.H))
                ops.philox_dropout_mask(mk, M * 2 * self.H, self.p_gru, seed, offset_dev, sid0 + l)
                masks[f'gru{l}'] = mk
        return masks

    def forward(self, poses, training, masks=None, save=True):
        """ConvDiscriminator.forward (multimodal_context_net.py:232-252): poses [B,34,27] -> sigmoid [B,1]."""
        ws = self.ws
        B, T0, D = poses.shape
        self.ctx = dict(B=B, T0=T0, D=D, masks=masks, poses=poses, training=training)
        x, tin, cin = poses, T0, D
        scale = shift = None
        self.Ts = [T0]
        self.ctx['fused_conv'] = self.fused_conv(B, T0, D)
        if self.ctx['fused_conv']:
            # the three convolutions and two train-mode BatchNorms in one 8-CTA-cluster launch (csrc/dconv_stack.cu)
            y0, y1, y2 = ws.get('d.y0', (B * 32, 16)), ws.get('d.y1', (B * 30, 8)), ws.get('d.y2', (B * 28, 8))
            st1, st2 = ws.get('d.st1', (64,)), ws.get('d.st2', (32,))
            bf = self.bufs
            ops.dconv_stack_fwd(poses, self.P('pre_conv.0.weight'), self.P('pre_conv.0.bias'), self.P('pre_conv.1.weight'), self.P('pre_conv.1.bias'),
                                bf['pre_conv.1.running_mean'], bf['pre_conv.1.running_var'], bf['pre_conv.1.num_batches_tracked'],
                                self.P('pre_conv.3.weight'), self.P('pre_conv.3.bias'), self.P('pre_conv.4.weight'), self.P('pre_conv.4.bias'),
                                bf['pre_conv.4.running_mean'], bf['pre_conv.4.running_var'], bf['pre_conv.4.num_batches_tracked'],
                                self.P('pre_conv.6.weight'), self.P('pre_conv.6.bias'), y0, y1, y2, st1, st2, B, T0, D, training, BN_EPS, BN_MOM)
            x, tin, cin = y2, 28, 8
            self.Ts = [34, 32, 30, 28]
        for li, (conv, bn) in enumerate(() if self.ctx['fused_conv'] else self.CONVS):
            w = self.P(conv + '.weight')
            cout, k = w.shape[0], w.shape[2]
            tout = tin - k + 1
            y = ws.get(f'd.y{li}', (B * tout, cout))
            ops.conv1d(x, w, self.P(conv + '.bias'), y, B=B, Tin=tin, Cin=cin, N=cout, k=k, pscale=scale, pshift=shift, pslope=1.0)
            if bn is not None:
                mean, rstd = ws.get(f'd.mean{li}', (cout,)), ws.get(f'd.rstd{li}', (cout,))
                scale, shift = ws.get(f'd.scale{li}', (cout,)), ws.get(f'd.shift{li}', (cout,))
                if training:
                    sums = ws.get(f'd.sums{li}', (2 * cout,), torch.float64); sums.zero_()
                    ops.col_stats(y, cout, B * tout, cout, sums)
                    ops.bn_finalize(sums, B * tout, cout, BN_EPS, BN_MOM, 1, self.P(bn + '.weight'), self.P(bn + '.bias'),
                                    self.bufs[bn + '.running_mean'], self.bufs[bn + '.running_var'], self.bufs[bn + '.num_batches_tracked'],
                                    mean, rstd, scale, shift)
                else:
                    ops.bn_eval_fold(self.P(bn + '.weight'), self.P(bn + '.bias'), self.bufs[bn + '.running_mean'],
                                     self.bufs[bn + '.running_var'], BN_EPS, None, scale, shift, cout)
            x, tin, cin = y, tout, cout
            self.Ts.append(tout)
        T = tin
        M = B * T
        H = self.H
        gmasks = [masks.get(f'gru{l}') for l in range(self.L)] if masks else None
        hsum = ws.get('d.hsum', (M, H)); o1 = ws.get('d.o1', (M, 1)); prob = ws.get('d.prob', (B, 1))
        if self.fused_stack(T, cin):
            # one launch for the 4-layer bidirectional GRU + both heads (csrc/dgru_stack.cu); writes exactly the buffers the per-layer
            # backward plan reads (layer outputs, saved gate planes, masked layer inputs, hsum, per-frame head output)
            L = self.L
            outs = [ws.get(f'd.out{l}', (M, 2 * H)) for l in range(L)]
            saved = [ws.get(f'd.saved{l}', (4, M, 2 * H)) for l in range(L)] if save else None
            drops = [(ws.get(f'd.drop{l}', (M, 2 * H)) if (gmasks is not None and l < L - 1 and gmasks[l] is not None) else None) for l in range(L)]
            mks = [(gmasks[l] if (gmasks is not None and l < L - 1) else None) for l in range(L)]
            a0 = self.arena.offsets['gru.weight_ih_l0']
            ops.dgru_stack_fwd(x, self.arena.flat[a0:], mks, outs, saved, M * 2 * H, drops, self.P('out.weight'), self.P('out.bias'),
                               self.P('out2.weight'), self.P('out2.bias'), hsum, o1, prob, B, T, cin, H, L, fast=config.fast())
            self.gru._fwd_inputs = x
        else:
            if getattr(self, '_gru_prep_stale', True):
                self.gru.prep()
                self._gru_prep_stale = False
            out = self.gru.forward(x, B, T, gmasks, save)
            ops.sum_halves(out, hsum, M, H)
            ops.linear(hsum, self.P('out.weight'), self.P('out.bias'), o1, M=M, K=H, N=1)
            ops.linear(o1, self.P('out2.weight'), self.P('out2.bias'), prob, M=B, K=T, N=1, act1=ops.ACT_SIGMOID)
        self.ctx['T'] = T
        return prob

    def fused_conv(self, B, T0, D):
        """The fused convolution-stack kernel covers the reference's ConvDiscriminator shapes (34 frames x 27 dims, 16 / 8 / 8 channels, k = 3)
        for up to 128 clips; TGB200_D_FUSED=0 keeps the per-operator plan."""
        m = self.m
        shapes = tuple(tuple(m.pre_conv[i].weight.shape) for i in (0, 3, 6))
        return config.d_fused() and B <= 128 and T0 == 34 and D == 27 and shapes == ((16, 27, 3), (8, 16, 3), (8, 8, 3))

    def fused_stack(self, T, I0):
        """The fused recurrent-stack kernel handles the discriminator's configuration (H = 64, <= 4 layers, <= 32 frames) when the GRU
        parameters are packed in the arena in gru_arena_order, which ParamArena guarantees; TGB200_D_FUSED=0 keeps the per-layer plan."""
        if not config.d_fused() or self.H != 64 or self.L > 4 or T > 32 or I0 > 32 or I0 % 4:
            return False
        names = []
        for l in range(self.L):
            for kind in ('weight_ih', 'bias_ih', 'weight_hh', 'bias_hh'):
                names += [f'gru.{kind}_l{l}', f'gru.{kind}_l{l}_reverse']
        return all(self.arena.adjacent(a, b) for a, b in zip(names, names[1:]))

    def backward(self, dlogit, need_dposes: bool, param_grads: bool = True):
        """dlogit [B,1] = d loss / d (pre-sigmoid output).  Returns d poses [B,34,27] if requested."""
        ws, c = self.ws, self.ctx
        B, T, H = c['B'], c['T'], self.H
        M = B * T
        masks = c['masks']
        gmasks = [masks.get(f'gru{l}') for l in range(self.L)] if masks else None
        if self.fused_stack(T, 8):
            # one launch: heads + 4 x [recurrence backward + data gradient through W_ih]; the recurrent layers' weight gradients follow on
            # the weight-gradient stream from the dgi / dgh it leaves behind
            L = self.L
            side.join(S_WGRAD)      # weight-gradient launches of an earlier backward may still be reading dgi / dgh
            outs = [ws[f'd.out{l}'] for l in range(L)]
            saved = [ws[f'd.saved{l}'] for l in range(L)]
            dgi = [ws.get(f'd.dgi{l}', (M, 6 * H)) for l in range(L)]
            dgh = [ws.get(f'd.dgh{l}', (M, 6 * H)) for l in range(L)]
            mks = [(gmasks[l] if (gmasks is not None and l < L - 1) else None) for l in range(L)]
            dy = ws.get('d.dxin', (M, 8))
            a0 = self.arena.offsets['gru.weight_ih_l0']
            ops.dgru_stack_bwd(dlogit, self.arena.flat[a0:], mks, outs, saved, M * 2 * H, ws['d.hsum'], ws['d.o1'], self.P('out.weight'),
                               self.P('out2.weight'), dgi, dgh, dy, self.G('out.weight'), self.G('out.bias'), self.G('out2.weight'),
                               self.G('out2.bias'), B, T, 8, H, L, fast=config.fast())
            for l in range(L - 1, -1, -1):
                if l == 0:
                    inp = ws['d.y2']
                elif gmasks is not None and gmasks[l - 1] is not None:
                    inp = ws[f'd.drop{l - 1}']
                else:
                    inp = ws[f'd.out{l - 1}']
                # all layers' gate gradients exist at once here (one launch produced them): the 12 small GEMMs go round-robin over three streams
                self.gru.weight_grads(l, inp, dgi[l], dgh[l], outs[l], B, T, stream=(S_WGRAD, S_WGRAD2, S_WGRAD3)[l % 3])
        else:
            do1 = ws.get('d.do1', (M, 1)); dhs = ws.get('d.dhs', (M, H)); dout = ws.get('d.dout', (M, 2 * H))
            ops.linear_wgrad(ws['d.o1'], dlogit, self.G('out2.weight'), self.G('out2.bias'), M=B, K=T, N=1)
            ops.linear_dgrad(dlogit, self.P('out2.weight'), do1, M=B, K=T, N=1)
            ops.linear_wgrad(ws['d.hsum'], do1, self.G('out.weight'), self.G('out.bias'), M=M, K=H, N=1)
            ops.linear_dgrad(do1, self.P('out.weight'), dhs, M=M, K=H, N=1)
            ops.dup_halves(dhs, dout, M, H)
            dy = self.gru.backward(dout, ws['d.y2'], B, 0, B, T, gmasks, True)
        dposes = None
        if c.get('fused_conv'):
            assert c['training'], 'backward through eval-mode BatchNorm is not on the hot path'
            dposes_buf = ws.get('d.da0', (B * 34, 27)) if need_dposes else None
            ops.dconv_stack_bwd(dy, c['poses'], ws['d.y0'], ws['d.y1'], ws['d.st1'], ws['d.st2'], self.P('pre_conv.0.weight'),
                                self.P('pre_conv.3.weight'), self.P('pre_conv.6.weight'), self.P('pre_conv.1.weight'), self.P('pre_conv.4.weight'),
                                self.G('pre_conv.0.weight'), self.G('pre_conv.0.bias'), self.G('pre_conv.3.weight'), self.G('pre_conv.3.bias'),
                                self.G('pre_conv.6.weight'), self.G('pre_conv.6.bias'), self.G('pre_conv.1.weight'), self.G('pre_conv.1.bias'),
                                self.G('pre_conv.4.weight'), self.G('pre_conv.4.bias'), dposes_buf, B, 34, 27)
            side.join(S_WGRAD)
            return dposes_buf.view(B, 34, 27) if need_dposes else None
        for li in (2, 1, 0):
            conv, _ = self.CONVS[li]
            w = self.P(conv + '.weight')
            cout, cin, k = w.shape
            tin, tout = self.Ts[li], self.Ts[li + 1]
            x = c['poses'] if li == 0 else ws[f'd.y{li - 1}']
            sc = ws[f'd.scale{li - 1}'] if li > 0 else None
            sh = ws[f'd.shift{li - 1}'] if li > 0 else None
            ops.conv1d_wgrad(x, dy, self.G(conv + '.weight'), self.G(conv + '.bias'), B=B, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k,
                             pscale=sc, pshift=sh, pslope=1.0)
            if li == 0 and not need_dposes:
                break
            da = ws.get(f'd.da{li}', (B * tin, cin))
            ops.conv1d_dgrad(dy, w, da, B=B, Tin=tin, Tout=tout, Cin=cin, N=cout, k=k)
            if li == 0:
                dposes = da.view(B, tin, cin)
                break
            bn = self.CONVS[li - 1][1]
            sums = ws.get(f'd.bsums{li - 1}', (2 * cin,), torch.float64); sums.zero_()
            mean, rstd = ws[f'd.mean{li - 1}'], ws[f'd.rstd{li - 1}']
            assert c['training'], 'backward through eval-mode BatchNorm is not on the hot path'
            ops.bn_bwd_reduce(da, x, B * tin, cin, mean, rstd, sc, sh, 1.0, sums)
            ops.bn_bwd_apply(da, x, da, B * tin, cin, mean, rstd, sc, sh, 1.0, self.P(bn + '.weight'), sums, self.G(bn + '.weight'),
                             self.G(bn + '.bias'))
            dy = da
        side.join(S_WGRAD)
        return dposes


# =====================================================================================================================
# EmbeddingNet (mode='pose'), eval mode: the feature extractor behind FGD
# =====================================================================================================================
class EmbeddingEngine:
    """PoseEncoderConv + PoseDecoderConv forward (embedding_net.py:42-82,165-217) with eval-mode BatchNorm folded into
    each producing GEMM's epilogue (scale/shift per output channel)."""

    def __init__(self, module):
        self.m = module
        self.ws: Optional[Workspace] = None

    def ensure(self, device):
        if self.ws is None or self.ws.device != device:
            self.ws = Workspace(device)
        self.p = {k: v.data for k, v in self.m.named_parameters()}
        self.b = dict(self.m.named_buffers())
        return self

    def _fold(self, tag, bn, conv_bias, C):
        sc, sh = self.ws.get(tag + '.sc', (C,)), self.ws.get(tag + '.sh', (C,))
        ops.bn_eval_fold(self.p[bn + '.weight'], self.p[bn + '.bias'], self.b[bn + '.running_mean'], self.b[bn + '.running_var'],
                         BN_EPS, conv_bias, sc, sh, C)
        return sc, sh

    def forward(self, poses, eps=None, variational=False, decode=True):
        """poses [B,T,D] -> (feat [B,32], mu, logvar, recon [B,T,D] or None)"""
        ws, p = self.ws, self.p
        B, T, D = poses.shape
        x, tin, cin = poses, T, D
        e = 'pose_encoder.'
        for i, (k, s) in enumerate(((3, 1), (3, 1), (4, 2))):
            w = p[f'{e}net.{i}.0.weight']
            cout = w.shape[0]
            sc, sh = self._fold(f'e{i}', f'{e}net.{i}.1', p[f'{e}net.{i}.0.bias'], cout)
            tout = _conv_out(tin, k, s)
            y = ws.get(f'emb.e{i}', (B * tout, cout))
            ops.conv1d(x, w, sh, y, B=B, Tin=tin, Cin=cin, N=cout, k=k, stride=s, escale=sc, act1=ops.ACT_LRELU, slope1=0.2)
            x, tin, cin = y, tout, cout
        w = p[e + 'net.3.weight']
        cout, k = w.shape[0], w.shape[2]
        tout = _conv_out(tin, k, 1)
        y = ws.get('emb.e3', (B * tout, cout))
        ops.conv1d(x, w, p[e + 'net.3.bias'], y, B=B, Tin=tin, Cin=cin, N=cout, k=k)
        # flatten is channel-major in the reference ([B,32,12] -> 384): read the Linear weight with (tap=t, chan=c) strides
        w0 = p[e + 'out_net.0.weight']
        n0 = w0.shape[0]
        assert w0.shape[1] == cout * tout, 'PoseEncoderConv.out_net is hard-wired to 34-frame clips (embedding_net.py:54-55)'
        sc, sh = self._fold('o0', e + 'out_net.1', p[e + 'out_net.0.bias'], n0)
        h0 = ws.get('emb.h0', (B, n0))
        ops.conv_gemm(y, w0, h0, B=B, Tin=tout, Tout=1, N=n0, Cin=cout, taps=tout, ldw=cout * tout, wsj=1, wsc=tout, escale=sc, bias=sh)
        w1 = p[e + 'out_net.3.weight']; n1 = w1.shape[0]
        sc, sh = self._fold('o1', e + 'out_net.4', p[e + 'out_net.3.bias'], n1)
        h1 = ws.get('emb.h1', (B, n1))
        ops.linear(h0, w1, sh, h1, M=B, K=n0, N=n1, escale=sc)
        w2 = p[e + 'out_net.6.weight']; n2 = w2.shape[0]
        h2 = ws.get('emb.h2', (B, n2))
        ops.linear(h1, w2, p[e + 'out_net.6.bias'], h2, M=B, K=n1, N=n2)
        mu = ws.get('emb.mu', (B, 32)); logvar = ws.get('emb.logvar', (B, 32))
        ops.linear(h2, p[e + 'fc_mu.weight'], p[e + 'fc_mu.bias'], mu, M=B, K=n2, N=32)
        ops.linear(h2, p[e + 'fc_logvar.weight'], p[e + 'fc_logvar.bias'], logvar, M=B, K=n2, N=32)
        feat = mu
        if variational:
            feat = ws.get('emb.z', (B, 32))
            ops.reparam_fwd(mu, logvar, eps, feat, B * 32)
        if not decode:
            return feat, mu, logvar, None
        d = 'decoder.'
        wp = p[d + 'pre_net.0.weight']; c0 = wp.shape[0]
        sc, sh = self._fold('d0', d + 'pre_net.1', p[d + 'pre_net.0.bias'], c0)
        g0 = ws.get('emb.g0', (B, c0))
        ops.linear(feat, wp, sh, g0, M=B, K=32, N=c0, escale=sc)
        wq = p[d + 'pre_net.3.weight']; c1 = wq.shape[0]
        g1 = ws.get('emb.g1', (B, c1))
        ops.linear(g0, wq, p[d + 'pre_net.3.bias'], g1, M=B, K=c0, N=c1)
        # view [B,4,L] channel-major -> read as channels-last through A strides (row stride 1, channel stride L)
        ch, L = 4, c1 // 4
        x, tin, cin = g1, L, ch
        a_kw = dict(lda=1, asc=L, a_bstride=c1)
        for i, idx in enumerate((0, 3)):                                   # two ConvTranspose1d(k=3) + BN + LeakyReLU(0.2)
            w = p[f'{d}net.{idx}.weight']                                  # [Cin, Cout, k]
            cout, k = w.shape[1], w.shape[2]
            sc, sh = self._fold(f't{i}', f'{d}net.{idx + 1}', p[f'{d}net.{idx}.bias'], cout)
            tout = tin + k - 1
            y = ws.get(f'emb.t{i}', (B * tout, cout))
            ops.conv_gemm(x, w, y, B=B, Tin=tin, Tout=tout, N=cout, Cin=cin, taps=k, dil=-1, pad=0, ldw=k, wsj=1, wsc=cout * k,
                          escale=sc, bias=sh, act1=ops.ACT_LRELU, slope1=0.2, **a_kw)
            x, tin, cin, a_kw = y, tout, cout, {}
        for i, idx in enumerate((6, 7)):
            w = p[f'{d}net.{idx}.weight']
            cout, k = w.shape[0], w.shape[2]
            tout = tin - k + 1
            y = ws.get(f'emb.c{i}', (B * tout, cout))
            ops.conv1d(x, w, p[f'{d}net.{idx}.bias'], y, B=B, Tin=tin, Cin=cin, N=cout, k=k)
            x, tin, cin = y, tout, cout
        return feat, mu, logvar, x.view(B, tin, cin)


# =====================================================================================================================
# WavEncoder / TextEncoderTCN called on their own (multimodal_context_net.py:25-28,57-61)
# =====================================================================================================================
class StandaloneEncoderEngine:
    """Forward of a stand-alone `WavEncoder()` / `TextEncoderTCN(...)` module (the reference constructs them directly in ContextEncoder,
    embedding_net.py:225-226, and nothing stops a user from calling them).  The generator's launch plans for the two encoders run
    unchanged on an arena over THIS module's own parameters: a parameter the plans call `audio_encoder.feat_extractor.0.weight` /
    `text_encoder.tcn...` is this module's `feat_extractor.0.weight` / `tcn...`.  Forward only: the result carries no autograd graph
    (training these modules goes through PoseGenerator / EmbeddingNet, whose engines own the hand-derived backward)."""
    WAV = GeneratorEngine.WAV
    wav_fast, wav_forward = GeneratorEngine.wav_fast, GeneratorEngine.wav_forward
    text_forward, make_masks = GeneratorEngine.text_forward, GeneratorEngine.make_masks
    _tcn_conv = staticmethod(GeneratorEngine._tcn_conv)

    def __init__(self, module, prefix):
        self.m, self.prefix = module, prefix
        self.arena = ParamArena(module)
        self.ws = None
        self.use_text = prefix == 'text_encoder.'
        self.use_audio = not self.use_text
        if self.use_text:
            self.E = module.embedding.weight.shape[1]
            self.H = module.tcn.network[0].conv1.weight_v.shape[0]
            self.n_tcn = len(module.tcn.network)
            self.tcn_k = module.tcn.network[0].conv1.weight_v.shape[2]
            self.p_emb, self.p_tcn = float(module.emb_dropout), float(module.tcn.network[0].dropout1.p)
        self.L, self.p_gru = 1, 0.0                                       # make_masks: no GRU masks

    def P(self, name):
        assert name.startswith(self.prefix), name
        return self.arena.params[name[len(self.prefix):]].data

    def ensure(self, device):
        if not self.arena.is_current() or self.ws is None or self.ws.device != device:
            self.arena.ensure(device)
            self.ws = Workspace(device)
        self.bufs = {self.prefix + k: v for k, v in self.m.named_buffers()}
        return self

    def run_wav(self, wav):
        B = wav.shape[0]
        feat = self.wav_forward(wav.detach().contiguous().float(), self.m.training, 1)
        return feat.view(B, -1, feat.shape[1]).clone()

    def run_text(self, ids, seed, offset_dev):
        B, T = ids.shape
        ws = self.ws
        for i in range(self.n_tcn):                                       # weight-normed filters, tap-major (+ per-tap transposes in fast mode)
            for j in (1, 2):
                q = f'text_encoder.tcn.network.{i}.conv{j}'
                v = self.P(q + '.weight_v')
                N, Cin, k = v.shape
                wT = ws.get(f'tcn.wT{i}_{j}', (k, Cin, N)) if config.fast() else None
                ops.weight_norm_fwd(v, self.P(q + '.weight_g'), ws.get(f'tcn.w{i}_{j}', (k, N, Cin)), wT, ws.get(f'tcn.inv{i}_{j}', (N,)), N, Cin, k)
        masks = self.make_masks(B, T, seed, offset_dev) if self.m.training else None
        feat = self.text_forward(ids.contiguous(), B, T, masks)
        return feat.view(B, T, feat.shape[1]).clone()
