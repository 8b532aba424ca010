"""train_iter_seq2seq / custom_loss - drop-in for scripts/train_eval/train_seq2seq.py:6-51.

Same signature and returned dict ({'loss': float}).  The whole iteration (forward, custom_loss, backward through the
33-step attention decoder and the packed encoder, clip_grad_norm_(5), Adam) is a fixed sequence of C-ABI kernel launches
with no host synchronisation until the single loss read-back; after two eager iterations the sequence is captured into
a CUDA graph per (batch shape, max text length) and replayed."""
from typing import Optional

import torch

from tgb200 import _lib, config, ops

_MAX_NORM = 5.0            # train_seq2seq.py:48
_GRAPH_WARMUP = 2
_injected_masks: Optional[dict] = None


def inject_masks(masks: Optional[dict]):
    """Tests: dropout keep-masks ({'enc0': [B*Tm,2H], 'dec0': [T,B,H]}) for the next call instead of Philox draws."""
    global _injected_masks
    _injected_masks = masks


def custom_loss(output, target, args, epoch):
    """train_seq2seq.py:6-36 on CUDA tensors, value only (the training path fuses value and gradient in tg_s2s_loss)."""
    _lib.require_cuda()
    B, T, D = output.shape
    loss = torch.zeros(1, dtype=torch.float64, device=output.device)
    scratch = torch.empty(T, B, D, device=output.device)
    ops.s2s_loss(output.contiguous().float(), target.contiguous().float(), loss, scratch, B, T, D, args.loss_regression_weight,
                 args.loss_kld_weight, args.loss_reg_weight)
    return loss[0].float()


class _Slot:
    def __init__(self):
        self.calls = 0
        self.graph = None
        self.failed = False
        self.static = None


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _enqueue(args, net, eng, optim, in_text, lens_dev, Tm, target, masks, loss_buf, host_step, world=1):
    dev = target.device
    B = target.shape[0]
    eng.arena.zero_grad()
    if masks is None and net.training:
        masks = eng.make_masks(B, Tm, net._noise.seed, net._noise.offset_dev(dev))
        net._noise.advance()
    eng.forward(in_text, lens_dev, Tm, target, True, masks, save=True)
    loss_buf.zero_()
    eng.loss_backward(target, float(args.loss_regression_weight), float(args.loss_kld_weight), float(args.loss_reg_weight), loss_buf)
    if world > 1:                # data parallel (train.py:93-96): sum the flat gradient arena over the ranks (NCCL; gloo in the CPU tests)
        from train_eval.train_gan import _allreduce_grads
        _allreduce_grads(eng.arena)
    eng.clip_and_step(optim, _MAX_NORM, host_step=host_step, world=world)


def train_iter_seq2seq(args, epoch, in_text, in_lengths, target_poses, net, optim):
    _lib.require_cuda()
    global _injected_masks
    masks, _injected_masks = _injected_masks, None
    net_ = net.module if isinstance(net, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else net
    if not target_poses.is_cuda and not _lib.TRACE_ONLY:
        raise _lib.TgError('train_iter_seq2seq runs on CUDA tensors only (sm_100a kernels, no CPU fallback)')
    assert net_.training, 'train_iter_seq2seq expects net.train() (BatchNorm batch statistics, train_seq2seq.py:39-51)'
    dev = target_poses.device
    B = target_poses.shape[0]
    target = target_poses.contiguous().float()
    lens_dev, Tm = net_.prepare_lengths(in_lengths, dev)
    eng = net_.engine().ensure(dev, 'train_%d_%d' % (B, Tm))
    ws = eng.ws
    loss_buf = ws.get('ti.loss', (1,), torch.float64)
    world = _world()
    use_graph = config.graphs() and masks is None and not _lib.TRACE_ONLY and torch.cuda.is_available() and world == 1   # the collective runs eagerly
    done = False
    if use_graph:
        g = optim.param_groups[0]
        key = (id(net_), id(optim), dev.index, B, Tm, tuple(in_text.shape), config.mode(), float(g['lr']), tuple(g['betas']),
               float(args.loss_regression_weight), float(args.loss_kld_weight), float(args.loss_reg_weight))
        key = key + (eng.arena.flat.data_ptr() if eng.arena.flat is not None else 0,)
        slots = eng.__dict__.setdefault('_graph_slots', {})          # lives and dies with the engine (no id() reuse across re-created models)
        for k in [k for k in slots if k[-1] != key[-1]]:
            del slots[k]
        slot = slots.setdefault(key, _Slot())
        slot.calls += 1
        if not slot.failed and slot.calls > _GRAPH_WARMUP and eng.arena.is_current():
            if slot.static is None:
                slot.static = dict(text=ws.get('ti.s_text', tuple(in_text.shape), torch.int64), lens=ws.get('ti.s_lens', (B,), torch.int64),
                                   target=ws.get('ti.s_target', tuple(target.shape)))
            st = slot.static
            st['text'].copy_(in_text); st['lens'].copy_(lens_dev); st['target'].copy_(target)
            if slot.graph is None:
                try:
                    eng.arena.bind_optimizer(optim)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        _enqueue(args, net_, eng, optim, st['text'], st['lens'], Tm, st['target'], None, loss_buf, host_step=False)
                    slot.graph = graph
                except Exception as exc:
                    slot.failed = True
                    import warnings
                    warnings.warn('tgb200: CUDA-graph capture of train_iter_seq2seq failed (%s); continuing with eager launches'
                                  % str(exc).splitlines()[0])
                    torch.cuda.synchronize()
            if slot.graph is not None:
                slot.graph.replay()
                eng.arena.note_steps(1)
                done = True
    if not done:
        _enqueue(args, net_, eng, optim, in_text.contiguous(), lens_dev, Tm, target, masks, loss_buf, host_step=True, world=world)
    return {'loss': float(loss_buf.cpu()[0])}
