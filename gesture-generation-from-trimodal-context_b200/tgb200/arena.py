"""Flat parameter / gradient / Adam-state arenas.

All trainable parameters of a module live in ONE contiguous fp32 buffer (each tensor 16-byte aligned); the
nn.Parameters keep their reference names and shapes but become views into it, and `.grad` becomes a view into a
second flat buffer.  That makes the optimiser one kernel launch (tg_adam_flat) and the data-parallel gradient
exchange a handful of large NCCL all-reduces instead of ~90 small ones."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops


def _ceil4(n: int) -> int:
    return (n + 3) & ~3


class ParamArena:
    def __init__(self, module: nn.Module, order: Optional[List[str]] = None, partial: bool = False):
        self.module = module
        self.partial = partial          # the bound optimiser may hold more parameters than this arena (one arena per sub-module)
        named = dict(module.named_parameters())
        names = list(order) if order is not None else list(named.keys())
        assert set(names) == set(named.keys()), (set(names) ^ set(named.keys()))
        self.names = [n for n in names if named[n].requires_grad]
        self.params: Dict[str, nn.Parameter] = {n: named[n] for n in names}
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in self.names:
            self.offsets[n] = off
            off += _ceil4(named[n].numel())
        self.numel = off
        self.flat = self.grad = self.exp_avg = self.exp_avg_sq = None
        self.step_dev = None
        self.step = 0
        self._bound_optim = None

    # -- storage ------------------------------------------------------------------------------------------
    def is_current(self) -> bool:
        if self.flat is None:
            return False
        base = self.flat.data_ptr()
        for n in (self.names[0], self.names[-1]):
            if self.params[n].data_ptr() != base + 4 * self.offsets[n]:
                return False
        return True

    def _grads_bound(self) -> bool:
        """Cheap per-step check (first / last parameter): optimizer.zero_grad() drops every .grad at once."""
        for n in (self.names[0], self.names[-1]):
            g = self.params[n].grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * self.offsets[n]:
                return False
        return True

    def bind_grads(self):
        """Points every .grad at its slice of the flat gradient buffer.  A parameter whose .grad was dropped - optimizer.zero_grad() sets
        it to None by default - means "gradient zero" to autograd, which would accumulate into a fresh tensor; its slice is therefore
        cleared before it is handed back (otherwise the previous iteration's gradient would be accumulated into)."""
        stale = [n for n in self.names if self.params[n].grad is None
                 or self.params[n].grad.data_ptr() != self.grad.data_ptr() + 4 * self.offsets[n]]
        if len(stale) == len(self.names):
            self.grad.zero_()                   # the usual case: every gradient was dropped - one memset
        for n in self.names:
            p = self.params[n]
            o, k = self.offsets[n], p.numel()
            view = self.grad[o:o + k].view(p.shape)
            if n in stale and len(stale) != len(self.names):
                view.zero_()
            p.grad = view

    def ensure(self, device) -> 'ParamArena':
        """(Re)builds the flat storage if the module was moved / re-created since the last call."""
        if self.is_current():
            if not self._grads_bound():
                self.bind_grads()
            return self
        old_m, old_v = self.exp_avg, self.exp_avg_sq
        self.flat = torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        for n in self.names:
            p = self.params[n]
            assert p.dtype == torch.float32, (n, p.dtype)
            o, k = self.offsets[n], p.numel()
            self.flat[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + k].view(p.shape)
            p.grad = self.grad[o:o + k].view(p.shape)
        if old_m is not None and old_m.numel() == self.numel:
            self.exp_avg, self.exp_avg_sq = old_m.to(device), old_v.to(device)
        else:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        if self.step_dev is None or self.step_dev.device != self.flat.device:
            self.step_dev = torch.full((1,), self.step, device=device, dtype=torch.int64)
        self._bound_optim = None
        return self

    def view(self, name: str) -> torch.Tensor:
        return self.params[name].data

    def gview(self, name: str) -> torch.Tensor:
        o = self.offsets[name]
        p = self.params[name]
        return self.grad[o:o + p.numel()].view(p.shape)

    def adjacent(self, a: str, b: str) -> bool:
        return self.offsets[b] == self.offsets[a] + self.params[a].numel()

    def zero_grad(self):
        self.grad.zero_()

    # -- optimiser ----------------------------------------------------------------------------------------
    def bind_optimizer(self, optim: torch.optim.Optimizer):
        """Makes a caller-constructed torch.optim.Adam (train.py:104-109) share our flat state so that its
        state_dict() stays meaningful while the update itself is one tg_adam_flat launch."""
        if self._bound_optim is optim:
            st = optim.state.get(self.params[self.names[0]], None)
            if st and 'exp_avg' in st and st['exp_avg'].data_ptr() == self.exp_avg.data_ptr() + 4 * self.offsets[self.names[0]]:
                return
            # optim.load_state_dict() (resume) replaced the state tensors: take the loaded moments / step count over below
        assert isinstance(optim, torch.optim.Adam), 'train_iter_gan expects the torch.optim.Adam objects of train.py:104-109'
        assert len(optim.param_groups) == 1
        g = optim.param_groups[0]
        assert g.get('weight_decay', 0) == 0 and not g.get('amsgrad', False) and not g.get('maximize', False)
        ours = {id(self.params[n]) for n in self.names}
        theirs = {id(p) for p in g['params'] if p.requires_grad}
        assert (ours <= theirs) if self.partial else (ours == theirs), 'optimizer parameters differ from the module parameters'
        steps = set()
        for n in self.names:
            p = self.params[n]
            st = optim.state.get(p, None)
            o, k = self.offsets[n], p.numel()
            if st and 'exp_avg' in st and st['exp_avg'].data_ptr() != self.exp_avg.data_ptr() + 4 * o:
                self.exp_avg[o:o + k].copy_(st['exp_avg'].reshape(-1))
                self.exp_avg_sq[o:o + k].copy_(st['exp_avg_sq'].reshape(-1))
                steps.add(int(st['step']))
        if steps:
            assert len(steps) == 1
            self.step = steps.pop()
            self.step_dev.fill_(self.step)
        self._step_tensor = torch.tensor(float(self.step))
        for n in self.names:
            p = self.params[n]
            o, k = self.offsets[n], p.numel()
            optim.state[p] = {'step': self._step_tensor, 'exp_avg': self.exp_avg[o:o + k].view(p.shape),
                              'exp_avg_sq': self.exp_avg_sq[o:o + k].view(p.shape)}
        self._bound_optim = optim

    def adam_step(self, optim: torch.optim.Optimizer, grad_scale: float = 1.0, host_step: bool = True):
        """One Adam update of every parameter.  host_step=False: the caller advances step_dev on the device
        (CUDA-graph replay) and calls note_steps() afterwards."""
        self.bind_optimizer(optim)
        g = optim.param_groups[0]
        if host_step:
            self.step += 1
            self._step_tensor.fill_(float(self.step))
        ops.increment_i64(self.step_dev, 1)
        b1, b2 = g['betas']
        ops.adam_flat(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.numel, float(g['lr']), float(b1), float(b2),
                      float(g['eps']), float(grad_scale), self.step_dev)

    def adam_begin(self, optim: torch.optim.Optimizer, host_step: bool = True):
        """First half of a SPLIT update: binds the optimiser and advances the step counters; follow with adam_range() calls that
        together cover [0, numel) exactly once."""
        self.bind_optimizer(optim)
        if host_step:
            self.step += 1
            self._step_tensor.fill_(float(self.step))
        ops.increment_i64(self.step_dev, 1)

    def adam_range(self, optim: torch.optim.Optimizer, lo: int, hi: int, grad_scale: float = 1.0):
        """Adam update of the flat range [lo, hi) (lo 16-byte aligned: a tensor boundary) at the step adam_begin() set."""
        if hi <= lo:
            return
        assert lo % 4 == 0
        g = optim.param_groups[0]
        b1, b2 = g['betas']
        ops.adam_flat(self.flat[lo:hi], self.grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], hi - lo, float(g['lr']), float(b1),
                      float(b2), float(g['eps']), float(grad_scale), self.step_dev)

    def note_steps(self, n: int):
        self.step += n
        if getattr(self, '_step_tensor', None) is not None:
            self._step_tensor.fill_(float(self.step))
