"""Side measurement (not a test): ms per auto-encoder training step (train_feature_extractor.train_iter, batch 128) with eager launches
and with CUDA-graph replay, and the fp32 CPU oracle beside it.  python tests/bench_ae.py > gpurun_out/bench_ae.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'), os.path.join(ROOT, 'tests')]


def main():
    import ae_checks
    import train_feature_extractor as tfx
    from oracle import embed_train_oracle as EO
    from oracle import synth
    from tgb200 import config, ops
    dev = torch.device('cuda:0')
    out = {}
    cfg = None
    for graphs in (False, True):
        old = config.set_graphs(graphs)
        cfg, net, opt = ae_checks.build(dev)
        net.train()
        tg = [synth.make_inputs(cfg, 128, seed=60 + i)['target'].to(dev) for i in range(4)]
        for i in range(6):
            tfx.train_iter(None, 0, tg[i % 4], net, opt)
        l0 = ops.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        n = 200
        for i in range(n):
            tfx.train_iter(None, 0, tg[i % 4], net, opt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out['graph' if graphs else 'eager'] = {'ms_per_step': ms, 'samples_per_s': 128e3 / ms, 'launches_per_step': (ops.launches() - l0) / n}
        config.set_graphs(old)
    sd = synth.embedding_net_state_dict(cfg)
    o = synth.zeros_like_opt(sd)
    tgt = synth.make_inputs(cfg, 128, seed=60)['target']
    EO.train_iter_ae_oracle(sd, o, 1, tgt, 5e-4, True)
    t0 = time.perf_counter()
    for s in range(10):
        EO.train_iter_ae_oracle(sd, o, 1, tgt, 5e-4, True)
    dt = (time.perf_counter() - t0) / 10
    out['cpu_oracle'] = {'ms_per_step': dt * 1e3, 'samples_per_s': 128 / dt, 'threads': torch.get_num_threads()}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
