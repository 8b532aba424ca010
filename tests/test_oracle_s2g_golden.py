"""CPU: the Speech2Gesture oracle (oracle/s2g_oracle.py) against the reference-executed golden (tests/golden/s2g_step.npz, written by
oracle/make_golden_s2g.py from the unmodified reference modules and train_iter_speech2gesture)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import s2g_oracle as SO
from oracle import synth
from oracle.make_golden import digest
from oracle.make_golden_s2g import B, D, D_LR_W, D_SEED, G_SEED, LR, N_PRE, T, W_GAN, W_REG, make_inputs


def _templates():
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gesture-generation-from-trimodal-context_b200')
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from model.speech2gesture import Discriminator, Generator
    return Generator(T, D, N_PRE), Discriminator(D)


def _close(a, ref, tol, what):
    a, ref = np.asarray(a), np.asarray(ref)
    scale = max(abs(ref[0]), 1e-9)
    assert abs(a[0] - ref[0]) <= tol * scale + 1e-7, (what, 'l2', a[0], ref[0])
    assert np.abs(a[2:] - ref[2:]).max() <= tol * max(np.abs(ref[2:]).max(), 1e-6) + 1e-6 * scale, what


def test_s2g_oracle_matches_reference_golden():
    torch.set_num_threads(8)
    g = np.load(os.path.join(GOLDEN, 's2g_step.npz'))
    Gm, Dm = _templates()
    gsd = synth.s2g_state_dict(Gm.state_dict(), G_SEED)
    dsd = synth.s2g_state_dict(Dm.state_dict(), D_SEED)
    spec, target = make_inputs(B, 7)
    out = SO.generator_forward(gsd, spec, target[:, :N_PRE], T, False)
    ref = torch.from_numpy(g['eval/out'])
    assert ((out - ref).norm() / ref.norm()).item() < 2e-5
    dis = SO.discriminator_forward(dsd, target, False)
    refd = torch.from_numpy(g['eval/dis'])
    assert ((dis - refd).norm() / refd.norm()).item() < 2e-5
    gopt, dopt = {}, {}
    for step in (1, 2):
        spec, target = torch.from_numpy(g[f's{step}/spec']), torch.from_numpy(g[f's{step}/target'])
        r = SO.train_iter_oracle(gsd, dsd, gopt, dopt, step, spec, target, N_PRE, W_REG, W_GAN, LR, LR * D_LR_W)
        for k, v in r['losses'].items():
            tol = 2e-5 if step == 1 else 1e-3      # step 2 starts from Adam's sign-like first update of fp32 round-off level gradients
            assert abs(v - float(g[f's{step}/loss/{k}'])) <= tol * abs(float(g[f's{step}/loss/{k}'])) + 1e-7, (step, k, v)
        if step == 1:                                   # (Adam's sign-like first step amplifies round-off in near-zero gradients afterwards)
            for k, gr in r['g_grads'].items():
                if g[f's{step}/ggrad/{k}'][0] < 1e-4:       # a bias in front of a train-mode BatchNorm: analytically zero, round-off only
                    assert digest(gr)[0] < 1e-4, k
                    continue
                _close(digest(gr), g[f's{step}/ggrad/{k}'], 5e-4, (step, 'ggrad', k))
        gsd, dsd, gopt, dopt = r['g_sd'], r['d_sd'], r['g_opt'], r['d_opt']
        for k in ('audio_encoder.first_net.1.1.running_var', 'decoder.0.1.running_mean', 'pre_pose_encoder.1.running_var'):
            _close(digest(gsd[k]), g[f's{step}/gpost/{k}'], 1e-4 if step == 1 else 1e-3, (step, 'gpost', k))
        for k in ('net.2.1.running_mean', 'net.3.1.running_var'):
            _close(digest(dsd[k]), g[f's{step}/dpost/{k}'], 1e-4 if step == 1 else 1e-3, (step, 'dpost', k))
