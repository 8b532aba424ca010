"""train_iter_gan - one adversarial G+D iteration, drop-in for scripts/train_eval/train_gan.py:13-103.

Same signature, same arithmetic, same returned dict of python floats ('loss', 'KLD', 'DIV_REG', 'gen', 'dis').
What differs is the schedule on the device (B200-first, results identical):
  * the generator weights do not change between the reference's three generator forwards of one iteration
    (train_gan.py:30,50,67; the D step in between only updates D), so the three forwards run as ONE pass over
    3*B clips - a third of the sequential GRU steps; each pass still draws its own dropout masks and noise, and
    BatchNorm running statistics receive the same number of momentum updates;
  * the WavEncoder features of that pass are computed once (the three passes see the same audio, same weights and
    train-mode batch statistics, hence bit-identical features);
  * every loss term and its gradient come from two fused kernels, the optimiser is one flat Adam launch per
    network, and the five logged scalars leave the device in a single copy (the reference syncs 7 times);
  * with torch.distributed initialised (one process per GPU), gradients are averaged with NCCL all-reduce over
    the flat gradient arena before Adam - the data-parallel equivalent of nn.DataParallel (train.py:93-96).
"""
import os
from typing import Dict, Optional

import torch

from tgb200 import _lib, config, ops
from tgb200.engine import S_DREAL, S_SPK, side


def add_noise(data):
    """train_gan.py:8-10 (unused by the reference: use_noisy_target=False at :17)."""
    return data + torch.randn_like(data) * 0.1


def _unwrap(m):
    return m.module if hasattr(m, 'module') and isinstance(m, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else m


class StepNoise:
    """Optional injection of every random draw of one iteration (tests): eps [n_pass][B,16], perm [B] int64,
    g_masks / d_masks lists of channels-last keep-mask dicts (or None)."""

    def __init__(self, eps=None, perm=None, g_masks=None, d_masks=None):
        self.eps, self.perm, self.g_masks, self.d_masks = eps, perm, g_masks, d_masks


_injected_noise: Optional[StepNoise] = None


def inject_noise(noise: Optional[StepNoise]):
    global _injected_noise
    _injected_noise = noise


def _dist_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _allreduce_grads(arena, lo=0, hi=None, bucket_floats=4 << 20):
    """Bucketed all-reduce (sum) of arena.grad[lo:hi], complete on return; averaging is folded into Adam's grad_scale."""
    for w in _allreduce_grads_async(arena, lo, hi, bucket_floats):
        w.wait()


def _allreduce_grads_async(arena, lo=0, hi=None, bucket_floats=4 << 20, ranges=None):
    """Launches the bucketed NCCL all-reduce (sum) of arena.grad[lo:hi] and returns the work handles (ranges: a list that receives the
    (start, end) of every bucket, in launch order)."""
    import torch.distributed as dist
    works = []
    hi = arena.numel if hi is None else hi
    for o in range(lo, hi, bucket_floats):
        works.append(dist.all_reduce(arena.grad[o:min(hi, o + bucket_floats)], op=dist.ReduceOp.SUM, async_op=True))
        if ranges is not None:
            ranges.append((o, min(hi, o + bucket_floats)))
    return works


class _Comm:
    """Gradient exchange of one iteration.  inline=True (eager launches, or the whole iteration captured into ONE CUDA graph with the
    NCCL kernels inside it): the recurrent layers' gradients are all-reduced as soon as they exist (`early`, called from the generator's
    backward on its weight-gradient stream) and the rest when the backward has finished (`finish`), so most of the exchange hides under
    the encoder backward.  inline=False (the iteration captured as graph segments split at the collectives): one exchange per network
    between two segments."""

    def __init__(self, inline):
        self.inline = inline
        self.pending = {}
        self.on_bucket = None         # callable(arena, lo, hi) run as soon as a bucket of the final exchange has been reduced

    def early(self, arena, lo, wait=False):
        if self.inline:
            works = _allreduce_grads_async(arena, lo=lo)
            self.pending[id(arena)] = (lo, works)
            if wait:                      # the caller's stream consumes the reduced range right away (early Adam of that range)
                for w in works:
                    w.wait()

    def finish(self, arena):
        lo, works = self.pending.pop(id(arena), (arena.numel, []))
        ranges = []
        rest = _allreduce_grads_async(arena, hi=lo, bucket_floats=(2 << 20) if self.on_bucket else (4 << 20), ranges=ranges)
        for w in works:
            w.wait()
        for w, (o, e) in zip(rest, ranges):          # pipelined: bucket k's consumer (Adam of that range) runs under bucket k+1's exchange
            w.wait()
            if self.on_bucket is not None:
                self.on_bucket(arena, o, e)


def _dp_sync_once(module, arena):
    """First data-parallel step of a (module, flat arena): rank 0's parameters and buffers become everybody's (what nn.DataParallel's
    replicate-from-device-0 does every forward, train.py:93-96), so identical initial weights do not depend on identical RNG seeds."""
    import torch.distributed as dist
    key = arena.flat.data_ptr()
    if getattr(arena, '_dp_synced', None) == key or (arena.flat.is_cuda and torch.cuda.is_current_stream_capturing()):
        return
    dist.broadcast(arena.flat, 0)
    for b in module.buffers():
        if b.is_floating_point():
            dist.broadcast(b, 0)
    arena._dp_synced = key


_CAPTURE_STREAMS = {}


def _capture_stream(device):
    """The iteration is captured on a HIGH-priority stream, the forked weight-gradient / bias / mask streams keep the default (lowest)
    priority: kernel nodes inherit it, so when the data-gradient GEMM that the next recurrence waits for and that layer's weight-gradient
    GEMMs are runnable at the same time, the block scheduler fills free SMs with the critical chain first."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1 if config.flat_prio() else max(-2, min(torch.cuda.Stream.priority_range())))
    return _CAPTURE_STREAMS[key]


class _GraphSlot:
    """One captured CUDA graph of the whole iteration for a fixed (modules, batch shape, schedule, hyper-parameters)."""

    def __init__(self):
        self.calls = 0
        self.graph = None
        self.failed = False
        self.static = None


def release_graphs(pose_decoder) -> int:
    """Drops every captured iteration graph of this generator (they are re-captured on demand).  With data parallelism the graph holds
    NCCL kernels of the process group's communicator: call this (and torch.cuda.synchronize()) BEFORE dist.destroy_process_group(),
    otherwise the communicator teardown waits on the graph forever (seen on 2 x B200, round-2 call L)."""
    ge = _unwrap(pose_decoder).engine()
    slots = ge.__dict__.get('_gan_graph_slots', {})
    n = len(slots)
    for slot in slots.values():
        slot.graph = None
    slots.clear()
    return n


_GRAPH_WARMUP = 2          # eager iterations (allocate every workspace) before the iteration is captured


def _flags(args, epoch):
    after = epoch > args.loss_warmup
    do_d = after and args.loss_gan_weight > 0.0
    z_type = getattr(args, 'z_type', 'speaker')
    do_div = z_type in ('speaker', 'random') and args.loss_reg_weight > 0.0
    do_kld = do_div and z_type == 'speaker'
    return after, do_d, do_div, do_kld


def train_iter_gan(args, epoch, in_text, in_audio, target_poses, vid_indices, pose_decoder, discriminator, pose_dec_optim, dis_optim):
    _lib.require_cuda()
    global _injected_noise
    noise, _injected_noise = _injected_noise, None
    G, D = _unwrap(pose_decoder), _unwrap(discriminator)
    dev = target_poses.device
    if not target_poses.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('train_iter_gan runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
    B, T, Dm = target_poses.shape
    after, do_d, do_div, do_kld = _flags(args, epoch)
    world = _dist_world()
    target = target_poses.contiguous().float()
    in_text = in_text.contiguous()
    in_audio = in_audio.contiguous().float()
    vid = vid_indices.contiguous() if vid_indices is not None else None

    use_graph = (config.graphs() and noise is None and not _lib.TRACE_ONLY and torch.cuda.is_available())
    sc = None
    if use_graph:
        gg, dg = pose_dec_optim.param_groups[0], dis_optim.param_groups[0]
        key = (id(G), id(D), id(pose_dec_optim), id(dis_optim), dev.index, B, T, Dm, tuple(in_audio.shape), after, do_d, do_div, do_kld,
               G.training, D.training, config.mode(), config.overlap(), world, float(gg['lr']), float(dg['lr']), tuple(gg['betas']), tuple(dg['betas']),
               float(args.loss_regression_weight), float(args.loss_gan_weight), float(args.loss_kld_weight), float(args.loss_reg_weight),
               int(args.n_pre_poses))
        ge0, de0 = G.engine(), D.engine()
        key = key + (ge0.arena.flat.data_ptr() if ge0.arena.flat is not None else 0, de0.arena.flat.data_ptr() if de0.arena.flat is not None else 0)
        slots = ge0.__dict__.setdefault('_gan_graph_slots', {})      # dies with the generator's engine: no stale graph can outlive its workspaces
        for k in [k for k in slots if k[-2:] != key[-2:]]:
            del slots[k]                                              # the arenas were rebuilt: graphs captured on the old storage are dead
        slot = slots.setdefault(key, _GraphSlot())
        slot.calls += 1
        if not slot.failed and slot.calls > _GRAPH_WARMUP and G.engine().arena.is_current() and D.engine().arena.is_current():
            sc = _run_graphed(slot, args, epoch, in_text, in_audio, target, vid, G, D, pose_dec_optim, dis_optim)
    if sc is None:
        sc = _enqueue_step(args, epoch, in_text, in_audio, target, vid, G, D, pose_dec_optim, dis_optim, noise, world, host_step=True)

    # ---- one device->host copy for the logged scalars (train_gan.py:94-102)
    s = sc if isinstance(sc, list) else sc.cpu().tolist()
    huber = s[0] / (B * T * Dm)
    ret: Dict[str, float] = {'loss': args.loss_regression_weight * huber}
    if do_kld:
        kld = -0.5 * s[2] / (B * 16)
        if kld:
            ret['KLD'] = args.loss_kld_weight * kld
    if do_div:
        div = s[1] / B
        if div:
            ret['DIV_REG'] = args.loss_reg_weight * div
    if do_d:
        ret['gen'] = args.loss_gan_weight * s[3]
        ret['dis'] = s[4] + s[5]
    return ret


def _stage(dst, src):
    """Copy into a graph's static input buffer.  Device sources go through an elementwise KERNEL, not cudaMemcpyAsync: a copy-engine
    D2D copy queues behind whatever host->device transfer a prefetcher has in flight (train_eval/staging.py) and would stall the step."""
    if src.is_cuda and src.dtype == dst.dtype and src.shape == dst.shape:
        torch.add(src, 0, out=dst)
    else:
        dst.copy_(src)


def _run_graphed(slot, args, epoch, in_text, in_audio, target, vid, G, D, g_opt, d_opt):
    """Replays (capturing on first use) the CUDA graph of one iteration on static input buffers.  Everything random is
    drawn inside the graph from device-resident Philox offsets and the Adam step counters live on the device, so every
    replay is a fresh iteration.  Returns the scalars buffer, or None if capture is not possible (falls back to eager)."""
    ge, de = G.engine(), D.engine()
    B = target.shape[0]
    ge.ensure(target.device, 'train_%d' % B)
    ws = ge.ws
    if slot.static is None:
        slot.static = dict(in_text=ws.get('ti.s_text', tuple(in_text.shape), torch.int64), in_audio=ws.get('ti.s_audio', tuple(in_audio.shape)),
                           target=ws.get('ti.s_target', tuple(target.shape)),
                           vid=ws.get('ti.s_vid', tuple(vid.shape), torch.int64) if vid is not None else None)
    st = slot.static
    _stage(st['in_text'], in_text); _stage(st['in_audio'], in_audio); _stage(st['target'], target)
    if vid is not None:
        _stage(st['vid'], vid)
    if slot.graph is None:
        try:
            ge.arena.bind_optimizer(g_opt); de.arena.bind_optimizer(d_opt)
            torch.cuda.synchronize()
            slot.host_sc = torch.zeros(8, dtype=torch.float64).pin_memory()
            slot.ev = torch.cuda.Event(external=True)          # an event-record NODE inside the graph: the host can wait on it after replay

            def early(sc_dev):
                # on a side branch of the graph: as a node of the main chain the 64-byte device-to-host copy held up the discriminator's
                # backward by ~36 us (profiles/r02_timeline_step.txt)
                from tgb200.engine import side, S_SCALARS
                with side.on(S_SCALARS):
                    slot.host_sc.copy_(sc_dev, non_blocking=True)
                    slot.ev.record()
            out = {'early': early}
            world = _dist_world()
            graphs, arenas = [], []
            if world > 1 and config.nccl_in_graph() and not getattr(slot, 'no_inline', False):
                # the whole iteration, NCCL kernels included, as ONE graph (possible since the recurrence kernels no longer spin on
                # grid-wide co-residency: a cluster kernel cannot dead-lock against a co-resident NCCL kernel)
                try:
                    comm = _Comm(inline=True)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=_capture_stream(target.device), capture_error_mode='thread_local'):
                        for arena in _step_segments(args, epoch, st['in_text'], st['in_audio'], st['target'], st['vid'], G, D, g_opt, d_opt, None,
                                                    world, False, out, comm):
                            comm.finish(arena)
                    graphs = [graph]
                except Exception as exc:
                    import warnings
                    warnings.warn('tgb200: capturing NCCL inside the CUDA graph failed (%s); falling back to graph segments split at the '
                                  'collectives' % (str(exc).splitlines()[0] if str(exc) else type(exc).__name__))
                    slot.no_inline = True
                    torch.cuda.synchronize()
                    out = {'early': early}
            if not graphs:
                comm = _Comm(inline=False)
                gen = _step_segments(args, epoch, st['in_text'], st['in_audio'], st['target'], st['vid'], G, D, g_opt, d_opt, None,
                                     world, False, out, comm)
                done = False
                while not done:                       # one CUDA graph per collective-free stretch of the iteration
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=_capture_stream(target.device)):
                        try:
                            arenas.append(next(gen))
                        except StopIteration:
                            done = True
                    graphs.append(graph)
            slot.graph, slot.arenas, slot.sc = graphs, arenas, out['sc']
        except Exception as exc:            # capture unsupported for some launch: stay on the eager path
            slot.failed = True
            import warnings
            root = exc
            while root.__context__ is not None:
                root = root.__context__
            warnings.warn('tgb200: CUDA-graph capture of train_iter_gan failed (%s | root cause: %s %s); continuing with eager launches'
                          % (str(exc).splitlines()[0], type(root).__name__, str(root).splitlines()[0] if str(root) else ''))
            torch.cuda.synchronize()
            return None
    for i, graph in enumerate(slot.graph):
        graph.replay()
        if i < len(slot.arenas):
            _Comm(inline=False).finish(slot.arenas[i])      # eager NCCL call between two graph segments
    _, do_d, _, _ = _flags(args, epoch)
    ge.arena.note_steps(1)
    if do_d:
        de.arena.note_steps(1)
    # The logged scalars (train_gan.py:94-102) are complete once the generator losses have been evaluated - before the generator backward,
    # its weight gradients and Adam run (~40 % of the step).  The graph copies them to pinned host memory at that point and records an
    # external event; waiting on IT instead of on the whole step lets the caller's next iteration (input staging, graph launch) be queued
    # while this one is still finishing, so the GPU never idles between iterations.  Stream order keeps every later read of the
    # parameters correct.
    slot.ev.synchronize()
    return slot.host_sc.tolist()


def _enqueue_step(args, epoch, in_text, in_audio, target, vid, G, D, pose_dec_optim, dis_optim, noise, world, host_step):
    """Eager iteration: enqueues every kernel (no host synchronisation) and runs the gradient all-reduces in line;
    returns the fp64 scalars buffer."""
    out = {}
    comm = _Comm(inline=True)
    for arena in _step_segments(args, epoch, in_text, in_audio, target, vid, G, D, pose_dec_optim, dis_optim, noise, world, host_step, out, comm):
        comm.finish(arena)
    return out['sc']


def _step_segments(args, epoch, in_text, in_audio, target, vid, G, D, pose_dec_optim, dis_optim, noise, world, host_step, result, comm=None):
    """Generator over the launch sequence of one iteration.  With world > 1 it yields the flat gradient arena at the two
    points where gradients must be summed across ranks (after D's backward, after G's backward): the caller performs
    the collective (NCCL) and resumes.  Each stretch between yields touches no collective, has all auxiliary streams
    joined at its ends and can therefore be captured as its own CUDA graph.  result['sc'] = fp64 scalars buffer."""
    dev = target.device
    B, T, Dm = target.shape
    after, do_d, do_div, do_kld = _flags(args, epoch)
    ge = G.engine().ensure(dev, 'train_%d' % B)
    de = D.engine().ensure(dev, 'train_%d' % B)
    ws = ge.ws
    if world > 1:
        _dp_sync_once(G, ge.arena); _dp_sync_once(D, de.arena)

    # ---- pass list: [D-step forward] + G-step forward + [permuted-speaker forward]
    passes = (['d'] if do_d else []) + ['g'] + (['r'] if do_div else [])
    n_pass = len(passes)
    Bt = n_pass * B
    ig = passes.index('g')
    off = G._noise.offset_dev(dev)
    seed = G._noise.seed
    # The text chain (embedding -> 8 dependent GEMMs) is the longest in front of the GRU: its masks and weight-normed filters are queued
    # first, on their own prioritised streams (engine.py: S_PREP), then the audio chain, then everything else of the iteration's top.
    g_masks = None
    if G.training and not (noise is not None and noise.g_masks is not None):
        g_masks = ge.make_masks(Bt, T, seed, off, split=True)
    ge.prep_weights('tcn')
    if ge.use_audio and (config.wav_first() or not ge.use_text):
        ge.start_wav(in_audio, G.training, n_pass)          # eager launches are host-bound: queue the audio chain early (side stream)
    pre_seq = ws.get('ti.pre', (B, T, Dm + 1))
    ops.make_pre_seq(target, pre_seq, B, T, Dm, args.n_pre_poses)

    vid_all = eps_all = None
    with side.on(S_SPK):        # noise / speaker ids feed the speaker branch that ge.forward queues on the same stream
        if ge.z_mode is not None:
            eps_all = ws.get('ti.eps', (Bt, 16))
            if noise is not None and noise.eps is not None:
                for i, e in enumerate(noise.eps[-n_pass:] if len(noise.eps) > n_pass else noise.eps):
                    eps_all[i * B:(i + 1) * B].copy_(e)
            else:
                ops.philox_normal(eps_all, Bt * 16, seed, off, 1000)
        if ge.z_mode == 'speaker':
            vid_all = ws.get('ti.vid', (Bt,), torch.int64)
            for i, p in enumerate(passes):
                if p == 'r':
                    perm = ws.get('ti.perm', (B,), torch.int64)
                    if noise is not None and noise.perm is not None:
                        perm.copy_(noise.perm)
                    else:
                        ops.philox_randperm(perm, B, seed, off, 1001)
                    ops.gather_i64(vid, perm, vid_all[i * B:(i + 1) * B], B)
                else:
                    vid_all[i * B:(i + 1) * B].copy_(vid)
    if G.training and noise is not None and noise.g_masks is not None:
        g_masks = _stack_masks(ws, noise.g_masks[-n_pass:] if len(noise.g_masks) > n_pass else noise.g_masks, B * T)

    d_drawn = {}
    if D.training and not (noise is not None and noise.d_masks is not None):
        # the masks of the two discriminator passes that follow the generator sweep, and the zeroing of the generator's gradient arena, are
        # queued now on the speaker stream (joined by ge.forward): as nodes of the main chain they sat between D's Adam and D(G(x))
        with side.on(S_SPK):
            for i in ((1, 2) if do_d else (2,)):
                d_drawn[i] = de.make_masks(B, T - 6, D._noise.seed, D._noise.offset_dev(dev), sid0=16 * i, tag=str(i))
    with side.on(S_SPK):
        ge.arena.zero_grad()

    def d_masks_for(i):
        if not D.training:
            return None
        if noise is not None and noise.d_masks is not None:
            return noise.d_masks[i]
        if i in d_drawn:
            return d_drawn[i]
        return de.make_masks(B, T - 6, D._noise.seed, D._noise.offset_dev(dev), sid0=16 * i, tag=str(i))

    sc = ws.get('ti.scalars', (8,), torch.float64)
    sc.zero_()
    dlogit = ws.get('ti.dlogit', (B, 1))
    if not do_d:
        de.prep_weights()
    if do_d:
        # D(real) forward + backward does not depend on the generator: it runs on an auxiliary stream underneath the
        # generator passes (train_gan.py:38,41-42; BatchNorm statistics still see real before fake)
        de.arena.zero_grad()
        with side.on(S_DREAL):
            masks_real = d_masks_for(0)

        def d_real():
            with side.on(S_DREAL):
                de.prep_weights()
                p_real = de.forward(target, D.training, masks_real)
                ops.bce_sigmoid(p_real, B, 1.0, 0.0, 1.0, sc[4:], dlogit)
                de.backward(dlogit, need_dposes=False)
        at = config.d_real_at()
        if at == 'top':
            d_real()
        else:
            # forked from inside the generator's forward (engine.py: beside_gru): at the GRU input, or behind the first recurrent layers
            ge.beside_gru, ge.beside_gru_at = d_real, {'concat': -1, 'gru0': 0, 'gru1': 1, 'gru2': 2, 'pre0': 100, 'pre1': 101, 'pre2': 102, 'pre3': 103}[at]

    # ---- all generator passes in one sweep
    ge.prep_weights('rest')
    poses, z, mu, logvar = ge.forward(pre_seq, in_text, in_audio, vid_all, eps_all, Bt, G.training, g_masks, n_bn_updates=n_pass)
    G._noise.advance()          # after the forward has joined the side-stream mask draws that still read this iteration's offset

    # ---- train D (train_gan.py:24-43)
    if do_d:
        side.join(S_DREAL)
        p_fake = de.forward(poses[0:B], D.training, d_masks_for(1))          # generator output of the D-step pass (detached)
        ops.bce_sigmoid(p_fake, B, -1.0, 1.0, 1.0, sc[5:], dlogit)
        de.backward(dlogit, need_dposes=False)
        if world > 1:
            yield de.arena
        de.arena.adam_step(dis_optim, grad_scale=1.0 / world, host_step=host_step)
        de.prep_weights()

    # ---- train G (train_gan.py:45-92); the generator's gradient arena was zeroed on a side stream at the top of the iteration
    out = poses[ig * B:(ig + 1) * B]
    p_gen = de.forward(out, D.training, d_masks_for(2))                       # runs in warm-up too (updates D's BN statistics)
    if D.training:
        D._noise.advance()
    d_out = ws.get('ti.dout', (B, T, Dm))
    dmu = ws.get('ti.dmu', (B, 16)); dlv = ws.get('ti.dlv', (B, 16))
    ir = passes.index('r') if do_div else None
    ops.gen_losses(out, target, poses[ir * B:(ir + 1) * B] if do_div else None,
                   z[ig * B:(ig + 1) * B] if do_div else None, z[ir * B:(ir + 1) * B] if do_div else None,
                   mu[ig * B:(ig + 1) * B] if do_kld else None, logvar[ig * B:(ig + 1) * B] if do_kld else None,
                   B, T * Dm, 16, float(args.loss_regression_weight), float(args.loss_reg_weight) if do_div else 0.0,
                   float(args.loss_kld_weight) if do_kld else 0.0, sc, d_out, dmu if do_kld else None, dlv if do_kld else None)
    ops.bce_sigmoid(p_gen, B, 1.0, 0.0, float(args.loss_gan_weight), sc[3:], dlogit)
    if result.get('early') is not None:
        result['early'](sc)          # every logged scalar is final here: the graph path reads them back before the generator backward runs
    if after:
        dposes = de.backward(dlogit, need_dposes=True)                        # D's own (stale) grads accumulate as in the reference
        ops.add(d_out, dposes, d_out, B * T * Dm)
    # The recurrent layers' parameters are the tail of the flat arena (22 of 53 MB) and their gradients are final as soon as the GRU's
    # backward has drained its weight-gradient stream: their all-reduce (data parallel) AND their Adam update run there and then, under the
    # text / audio encoder backward; only the encoders' share of Adam is left for the end of the iteration.
    gru_lo = ge.arena.offsets.get('gru.weight_ih_l0')
    split = gru_lo is not None and gru_lo % 4 == 0 and (world == 1 or (comm is not None and comm.inline))
    if split:
        ge.arena.adam_begin(pose_dec_optim, host_step=host_step)

        def on_gru_grads():
            if world > 1:
                comm.early(ge.arena, gru_lo, wait=True)
            ge.arena.adam_range(pose_dec_optim, gru_lo, ge.arena.numel, grad_scale=1.0 / world)
        ge.on_gru_grads = on_gru_grads
    else:
        ge.on_gru_grads = None
    ge.backward(d_out, ig * B, (ig + 1) * B, d_mu=dmu if do_kld else None, d_logvar=dlv if do_kld else None)
    ge.on_gru_grads = None
    # the scalar read-back branch rejoins only here (as a member of the weight-gradient join group it made the discriminator's backward wait
    # ~27 us for the device-to-host copy: profiles/r02_timeline_step_3p78ms.txt, 2072-2099 us)
    from tgb200.engine import S_SCALARS
    side.join(S_SCALARS)
    done = []
    if world > 1:
        if split and os.environ.get('TGB200_DP_PIPELINE', '0') == '1':
            # opt-in: 8 MB buckets with Adam of bucket k under the exchange of bucket k+1.  Measured on 2 x B200: 4.38 ms/step against 4.32
            # with one wait and one Adam launch for the whole range (four more NCCL launches cost more than the overlap returns)
            def on_bucket(arena, o, e):
                if arena is ge.arena:
                    ge.arena.adam_range(pose_dec_optim, o, e, grad_scale=1.0 / world)
                    done.append((o, e))
            comm.on_bucket = on_bucket
        yield ge.arena
        if comm is not None:
            comm.on_bucket = None
    if split and done:
        assert done[0][0] == 0 and done[-1][1] == gru_lo and all(a[1] == b[0] for a, b in zip(done, done[1:])), done
    elif split:
        ge.arena.adam_range(pose_dec_optim, 0, gru_lo, grad_scale=1.0 / world)
    else:
        ge.arena.adam_step(pose_dec_optim, grad_scale=1.0 / world, host_step=host_step)
    result['sc'] = sc


def _stack_masks(ws, per_pass, rows_per_pass):
    """Concatenates per-pass mask dicts ([B*T, C] each) into one dict of [n_pass*B*T, C] tensors."""
    out = {}
    keys = set()
    for m in per_pass:
        if m:
            keys |= set(m.keys())
    for k in keys:
        ref = next(m[k] for m in per_pass if m and k in m)
        C = ref.shape[-1]
        buf = ws.get('ti.mask.' + k, (len(per_pass) * rows_per_pass, C))
        for i, m in enumerate(per_pass):
            if m and k in m:
                buf[i * rows_per_pass:(i + 1) * rows_per_pass].copy_(m[k].reshape(rows_per_pass, C))
            else:
                buf[i * rows_per_pass:(i + 1) * rows_per_pass].fill_(1.0)
        out[k] = buf
    return out
