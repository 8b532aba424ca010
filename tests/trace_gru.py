"""Development aid (not a test): duration and per-step phase timeline of the tensor-core GRU recurrence kernels at the benchmark
shape.  TGB200_GRU_LEGACY=1 selects the L2-counter-stepped kernels, default = the cluster / multicast kernels.
Cluster-kernel stamp slots (CTA 0 of cluster (0,0)): 1 accumulator ready, 2 gate scratch staged, 3 gates done, 4 cluster arrive,
5 cluster wait returned, 6 first operand group landed (MMA thread), 7 MMAs issued + commit."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import torch
from tgb200 import _lib, ops
dev = torch.device('cuda:0')
lib = _lib.load()


def run(kind, B, T=34, H=300, reps=20, trace_steps=(1, 2, 3, 16, 32)):
    M = B * T
    g = torch.Generator().manual_seed(0)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev)
    gi = r(M, 6 * H); whh = [r(3 * H, H), r(3 * H, H)]; bhh = [r(3 * H), r(3 * H)]
    out = torch.zeros(M, 2 * H, device=dev); saved = torch.zeros(4, M, 2 * H, device=dev)
    sync = torch.zeros(max(ops.gru_tf32_sync_ints(B, H), 64), dtype=torch.int32, device=dev)
    trace = torch.zeros(16 * T, dtype=torch.int64, device=dev)
    dout = r(M, 2 * H); dgi = torch.empty(M, 6 * H, device=dev); dgh = torch.empty(M, 6 * H, device=dev)
    partial = torch.empty(max(ops.gru_bwd_tf32_scratch_floats(B, H), 1), device=dev)
    whhT = [w.t().contiguous() for w in whh]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def launch():
        if kind == 'fwd':
            ops.gru_layer_fwd_tf32(gi, whh[0], whh[1], bhh[0], bhh[1], out, saved, M * 2 * H, sync, B, T, H)
        else:
            ops.gru_layer_bwd_tf32(dout, out, saved[0], M * 2 * H, whhT[0], whhT[1], dgi, dgh, partial, sync, B, T, H)
    if kind == 'bwd':
        ops.gru_layer_fwd_tf32(gi, whh[0], whh[1], bhh[0], bhh[1], out, saved, M * 2 * H, sync, B, T, H)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    times.sort()
    print('%s B=%d T=%d H=%d legacy=%s: median %.1f us  min %.1f us  (%.2f us/step)' % (
        kind, B, T, H, os.environ.get('TGB200_GRU_LEGACY', '0'), times[len(times) // 2], times[0], times[len(times) // 2] / T), flush=True)
    lib.tg_debug_gru_trace(ctypes.c_void_p(trace.data_ptr()))
    launch(); torch.cuda.synchronize()
    lib.tg_debug_gru_trace(ctypes.c_void_p(0))
    tr = trace.cpu().view(T, 16).numpy()
    for s in trace_steps:
        if s >= T:
            continue
        row = tr[s]
        if row.max() == 0:
            continue
        t0 = row[row > 0].min()
        print('  step %2d |' % s, ' '.join('%d:%5.2f' % (i, (v - t0) / 1e3) for i, v in enumerate(row) if v > 0),
              '| since prev step slot3: %.2f us' % ((row[3] - tr[s - 1][3]) / 1e3 if tr[s - 1][3] > 0 and row[3] > 0 else float('nan')))


if __name__ == '__main__':
    print('resident 8-CTA clusters: %d; batch tiles fwd*1000+bwd: B=384 %d  B=128 %d' % (
        lib.tg_debug_gru_cluster_occupancy(300, 16), lib.tg_debug_gru_cluster_tiles(384, 300), lib.tg_debug_gru_cluster_tiles(128, 300)), flush=True)
    run('fwd', 384); run('fwd', 128); run('bwd', 128)
