"""CPU, world_size 2, gloo: the data-parallel host logic of train_iter_gan - bucketed all-reduce of the flat gradient
arena (one process per GPU in production, NCCL over NVLink; here the same code path over gloo) and the sharded FGD
statistics reduction.  No kernel is launched."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gpu_util import build_ours
        from oracle.make_golden import golden_cfg
        from tgb200.engine import gru_arena_order
        from tgb200.arena import ParamArena
        from train_eval.train_gan import _allreduce_grads, _dist_world
        assert _dist_world() == world
        cfg = golden_cfg()
        args, G, D, _, _ = build_ours(cfg, None)
        arena = ParamArena(D, gru_arena_order([n for n, _ in D.named_parameters()])).ensure(torch.device('cpu'))
        g = torch.Generator().manual_seed(100 + rank)
        arena.grad.copy_(torch.randn(arena.numel, generator=g))
        mine = arena.grad.clone()
        _allreduce_grads(arena, bucket_floats=10007)          # several ragged buckets
        others = [torch.randn(arena.numel, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        expect = sum(others)
        ok = torch.allclose(arena.grad, expect, atol=1e-5) and torch.allclose(others[rank], mine)
        # .grad views of the parameters see the reduced values (they alias the flat buffer)
        p = dict(D.named_parameters())['gru.weight_hh_l2']
        o = arena.offsets['gru.weight_hh_l2']
        ok = ok and torch.equal(p.grad.reshape(-1), arena.grad[o:o + p.numel()])
        # FGD: per-rank sufficient statistics -> all-reduce -> identical scores on every rank
        from model.embedding_space_evaluator import EmbeddingSpaceEvaluator
        ev = EmbeddingSpaceEvaluator.__new__(EmbeddingSpaceEvaluator)
        ev.device = torch.device('cpu'); ev.reset()
        feats_r = torch.randn(64, 32, generator=torch.Generator().manual_seed(7 + rank)).double()
        feats_g = (torch.randn(64, 32, generator=torch.Generator().manual_seed(17 + rank)) * 1.3 + 0.2).double()
        for acc, f in ((ev.acc_real, feats_r), (ev.acc_gen, feats_g)):
            acc[0] = f.shape[0]; acc[1:33] = f.sum(0); acc[33:] = (f.t() @ f).reshape(-1)
        ev.acc_misc[0] = (feats_r - feats_g).abs().sum()
        fgd, fdist = ev.get_scores()
        q.put((rank, bool(ok), float(fgd), float(fdist)))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_and_fgd_reduction_world2():
    import numpy as np
    from oracle import trimodal_oracle as O
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert abs(res[0][2] - res[1][2]) < 1e-9 and abs(res[0][3] - res[1][3]) < 1e-9
    # the reduced FGD equals the oracle's FGD over the union of both ranks' features
    fr = np.vstack([torch.randn(64, 32, generator=torch.Generator().manual_seed(7 + r)).double().numpy() for r in range(2)])
    fg = np.vstack([(torch.randn(64, 32, generator=torch.Generator().manual_seed(17 + r)) * 1.3 + 0.2).double().numpy() for r in range(2)])
    fgd, fdist = O.fgd_scores(fg, fr)
    assert abs(res[0][2] - fgd) <= 1e-6 * abs(fgd) and abs(res[0][3] - fdist) <= 1e-9 * abs(fdist)
