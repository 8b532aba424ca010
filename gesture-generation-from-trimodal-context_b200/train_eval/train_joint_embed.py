"""train_iter_embed / eval_embed - drop-in for scripts/train_eval/train_joint_embed.py:5-65: the pose auto-encoder
(EmbeddingNet(mode='pose'), the FGD feature extractor) and the joint-embedding model (EmbeddingNet(mode='random'): ContextEncoder +
PoseEncoderConv + PoseDecoderGRU, `joint_step` below).  Same signatures, same returned values.

One optimiser step = the train-mode forward, the L1 reconstruction loss (value + gradient in one kernel), the hand-derived
backward and one flat Adam launch (tgb200.embed_engine.AutoEncoderTrainEngine): 82 launches of a few microseconds each, no host
synchronisation until the single loss read-back.  After two eager steps the sequence is captured into a CUDA graph per
(net, optimiser, batch shape) and replayed - the step is launch-latency bound, so this is where the time goes.
train_feature_extractor.train_iter (which adds the frame-difference term) shares `ae_step`."""
import random

import torch

from tgb200 import _lib, config

_GRAPH_WARMUP = 2          # eager steps (allocate every workspace) before the step is captured


class _Slot:
    def __init__(self):
        self.calls = 0
        self.graph = None
        self.failed = False
        self.static = None


def _unwrap(m):
    return m.module if isinstance(m, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else m


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _allreduce(arena):
    """Data parallel (one process per GPU, train.py:93-96 wraps every model family in nn.DataParallel): sum the flat gradient arena over
    the ranks (NCCL; gloo in the CPU tests); the mean is taken inside Adam (grad_scale = 1/world).  BatchNorm statistics stay per rank,
    like DataParallel's per-replica statistics."""
    from train_eval.train_gan import _allreduce_grads
    _allreduce_grads(arena)


def _enqueue(eng, optim, target, use_diff, weight, acc, host_step, world=1):
    eng.arena.zero_grad()                                   # optim.zero_grad(), train_joint_embed.py:9
    _, _, recon = eng.forward(target, training=True)        # :17-19
    acc.zero_()
    d_rec = eng.loss(recon, target, use_diff, weight, acc)  # :21-29,46
    eng.backward(d_rec)                                     # :48
    if world > 1:
        _allreduce(eng.arena)
    eng.arena.adam_step(optim, 1.0 / world, host_step=host_step)    # :49


def ae_step(net, optim, target_data, use_diff: bool, weight: float = 1.0) -> float:
    """One auto-encoder step on target_data [B,34,27]; returns recon_loss (python float, the step's only host read-back)."""
    _lib.require_cuda()
    net_ = _unwrap(net)
    if getattr(net_, 'mode', None) != 'pose':
        raise ValueError("ae_step trains the pose auto-encoder (EmbeddingNet(mode='pose')); the joint-embedding model goes through joint_step")
    if not target_data.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('the auto-encoder step runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
    assert net_.training, 'the auto-encoder step expects net.train() (BatchNorm batch statistics)'
    dev = target_data.device
    target = target_data.detach().contiguous().float()
    eng = net_.train_engine().ensure(dev)
    acc = eng.ws.get('ae.acc', (2,), torch.float64)
    world = _world()
    use_graph = config.graphs() and not _lib.TRACE_ONLY and torch.cuda.is_available() and world == 1    # the collective runs eagerly
    done = False
    if use_graph:
        g = optim.param_groups[0]
        key = (id(optim), dev.index, tuple(target.shape), bool(use_diff), float(weight), float(g['lr']), tuple(g['betas']))
        slot = eng.graph_slots.setdefault(key, _Slot())          # slots live on the engine: they die with the net they captured
        slot.calls += 1
        if not slot.failed and slot.calls > _GRAPH_WARMUP and eng.arena.is_current():
            if slot.static is None:
                slot.static = eng.ws.get('ae.s_target', tuple(target.shape))
            slot.static.copy_(target)
            if slot.graph is None:
                try:
                    eng.arena.bind_optimizer(optim)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        _enqueue(eng, optim, slot.static, use_diff, weight, acc, host_step=False)
                    slot.graph = graph
                except Exception as exc:                    # capture unsupported for some launch: stay on the eager path
                    slot.failed = True
                    import warnings
                    warnings.warn('tgb200: CUDA-graph capture of the auto-encoder step failed (%s); continuing with eager launches'
                                  % str(exc).splitlines()[0])
                    torch.cuda.synchronize()
            if slot.graph is not None:
                slot.graph.replay()
                eng.arena.note_steps(1)
                done = True
    if not done:
        _enqueue(eng, optim, target, use_diff, weight, acc, host_step=True, world=world)
    return float(acc.cpu()[0])


def _resolve_mode(net_, mode):
    """input_mode of EmbeddingNet.forward (embedding_net.py:277-296): None -> net.mode; 'random' flips a Python coin."""
    if mode is None:
        mode = net_.mode
    if mode == 'random':
        mode = 'speech' if random.random() > 0.5 else 'pose'
        if _world() > 1:                                                        # every rank must decode (and all-reduce) the same branch
            import torch.distributed as dist
            pick = [mode]
            dist.broadcast_object_list(pick, src=0)
            mode = pick[0]
    assert mode in ('speech', 'pose'), mode
    return mode


def joint_step(args, net, optim, in_text, in_audio, target_data, mode) -> float:
    """One step of the joint-embedding model (train_joint_embed.py:5-51, variational_encoding=False): forward of both encoders, decode
    the latent `mode` resolves to, loss = sum_b mean|recon - target|, backward through the decoder and THAT encoder, Adam on the
    parameters that received a gradient (torch.optim.Adam skips the others).  Eager launches (no CUDA graph: the branch changes
    from step to step)."""
    _lib.require_cuda()
    net_ = _unwrap(net)
    if not target_data.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('the joint-embedding step runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
    assert net_.training, 'train_iter_embed expects net.train()'
    branch = _resolve_mode(net_, mode)
    target = target_data.detach().contiguous().float()
    eng = net_.joint_engine().ensure(target.device)
    for a in eng.arenas():
        a.zero_grad()                                                           # optim.zero_grad(), :9
    pre_seq = target[:, 0:args.n_pre_poses]                                     # :6
    r = eng.forward(in_text, in_audio, pre_seq, target, branch, True)           # :17-19
    acc = eng.ws.get('jd.acc', (2,), torch.float64)
    acc.zero_()
    eng.backward(eng.loss(r['out'], target, acc))                               # :21-29,46-48
    world = _world()
    for a in ((eng.a_ctx if branch == 'speech' else eng.a_pose), eng.a_dec):    # :49
        if world > 1:
            _allreduce(a)
        a.adam_step(optim, 1.0 / world, host_step=True)
    return float(acc.cpu()[0])


def train_iter_embed(args, epoch, in_text, in_audio, target_data, net, optim, mode=None):
    """train_joint_embed.py:5-51 with variational_encoding=False (:12-15): loss = sum over the batch of the per-sample mean L1
    (the frame-difference term is switched off there, :24).  A mode='pose' net ignores in_text / in_audio (context_encoder is None,
    embedding_net.py:282) and mode must resolve to 'pose'; a joint-embedding net decodes the branch `mode` resolves to."""
    if getattr(_unwrap(net), 'context_encoder', None) is not None:
        return {'loss': joint_step(args, net, optim, in_text, in_audio, target_data, mode)}
    assert mode in (None, 'pose'), "EmbeddingNet(mode='pose') has no context encoder: input_mode must be 'pose' (embedding_net.py:295-303)"
    return {'loss': ae_step(net, optim, target_data, use_diff=False)}


def eval_embed(in_text, in_audio, pre_poses, target_poses, net, mode=None):
    """train_joint_embed.py:54-65 -> (loss, recon_poses): batch mean of the per-sample mean L1, whatever mode (train / eval) the
    net is in - like the reference, a train-mode net normalises with batch statistics and updates its running statistics."""
    _lib.require_cuda()
    net_ = _unwrap(net)
    if not target_poses.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('eval_embed runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
    target = target_poses.detach().contiguous().float()
    if getattr(net_, 'context_encoder', None) is not None:
        eng = net_.joint_engine().ensure(target.device)
        r = eng.forward(in_text, in_audio, pre_poses, target, _resolve_mode(net_, mode), net_.training)
        acc = eng.ws.get('jd.acc_eval', (2,), torch.float64)
        acc.zero_()
        eng.loss(r['out'], target, acc, want_grad=False)
        return (acc[1] / target.shape[0]).float(), r['out'].clone()
    assert mode in (None, 'pose')
    eng = net_.train_engine().ensure(target.device)
    _, _, recon = eng.forward(target, training=net_.training)
    acc = eng.ws.get('ae.acc_eval', (2,), torch.float64)
    acc.zero_()
    eng.loss(recon, target, False, 1.0, acc, want_grad=False)
    loss = (acc[1] / target.shape[0]).float()
    return loss, recon.clone()
