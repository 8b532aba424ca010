"""Side measurement (not a test): ms per joint-embedding training step (train_iter_embed(mode='random'), batch 128, eager launches) in both
arithmetic modes, launches per step, and the fp32 CPU oracle beside it.  python tests/bench_joint.py > gpurun_out/bench_joint.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'), os.path.join(ROOT, 'tests')]


def main(dry=False):
    import contextlib
    import joint_checks
    from oracle import joint_embed_oracle as J
    from oracle import synth
    from oracle import trimodal_oracle as O
    from tgb200 import config, ops
    from train_eval.train_joint_embed import train_iter_embed
    # --dry: the same code on CPU tensors over the NumPy C-ABI emulator with a toy size (checks this script, measures nothing)
    dev = torch.device('cpu' if dry else 'cuda:0')
    cfg = O.HotPathConfig(n_words=300 if dry else 20000, n_speakers=12 if dry else 1371)
    B, n, warm = (2, 2, 1) if dry else (128, 40, 6)
    out = {}
    stack = contextlib.ExitStack()
    if dry:
        import cabi_emulator
        stack.enter_context(cabi_emulator.installed())
        config.set_graphs(False)
    for mode in ('tf32', 'fp32'):
        old = config.set_mode(mode)
        _, args, net, opt = joint_checks.build(dev, cfg)
        net.train()
        data = [{k: v.to(dev) for k, v in synth.make_inputs(cfg, B, seed=60 + i).items()} for i in range(4)]
        f = lambda i: train_iter_embed(args, 0, data[i % 4]['in_text'], data[i % 4]['in_audio'], data[i % 4]['target'], net, opt, mode='random')
        for i in range(warm):
            f(i)
        l0 = ops.launches()
        if dry:
            t0 = time.perf_counter()
            for i in range(n):
                f(i)
            ms = (time.perf_counter() - t0) * 1e3 / n
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for i in range(n):
                f(i)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
        out[mode] = {'ms_per_step': ms, 'samples_per_s': B * 1e3 / ms, 'launches_per_step': (ops.launches() - l0) / n}
        config.set_mode(old)
    stack.close()
    sd = synth.joint_embedding_state_dict(cfg)
    nb = 2 if dry else 16
    d = synth.make_inputs(cfg, nb, seed=60)
    eps = torch.zeros(nb, 32)
    t0 = time.perf_counter()
    J.train_iter_embed_oracle(sd, synth.zeros_like_opt(sd), {}, d['in_text'], d['in_audio'], d['target'], cfg.n_pre_poses, 'speech', eps, 5e-4)
    dt = time.perf_counter() - t0
    out['cpu_oracle'] = {'ms_per_step': dt * 1e3, 'batch': nb, 'samples_per_s': nb / dt, 'threads': torch.get_num_threads()}
    print(json.dumps(out))


if __name__ == '__main__':
    main(dry='--dry' in sys.argv)
