"""Long-form inference driver (scripts/synthesize.py:36-209): host logic vs. fixtures produced by EXECUTING the
reference's own generate_gestures around a deterministic stub generator (oracle/make_golden_synthesize.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.make_golden_synthesize import golden_args
from oracle.synthesize_stub import StubGenerator, StubSeq2Seq, StubVocab, make_clip

CASES = (('short', 1.7, False), ('long', 9.3, False), ('fade', 6.1, True))


@pytest.mark.parametrize('tag,seconds,fade', CASES)
def test_driver_matches_reference_execution(tag, seconds, fade):
    from synthesize import generate_gestures
    g = np.load(os.path.join(GOLDEN, 'synthesize_driver.npz'))
    audio, words, seed = make_clip(seconds, seed=len(tag))
    out = generate_gestures(golden_args(), StubGenerator(), StubVocab(), audio, words, vid=7, seed_seq=seed, fade_out=fade)
    assert out.shape == g[tag].shape
    assert np.abs(out - g[tag]).max() < 2e-6, np.abs(out - g[tag]).max()


@pytest.mark.parametrize('tag,seconds,fade', (('s2s_short', 1.7, False), ('s2s_long', 9.3, False), ('s2s_fade', 6.1, True)))
def test_seq2seq_driver_matches_reference_execution(tag, seconds, fade):
    """args.model == 'seq2seq': ragged word-id input, seed hand-off, cross-fade AND the seq2seq-only cubic smoothing of every window join
    (synthesize.py:134-136,163-185) vs the executed reference driver."""
    from synthesize import generate_gestures
    g = np.load(os.path.join(GOLDEN, 'synthesize_driver.npz'))
    audio, words, seed = make_clip(seconds, seed=len(tag))
    args = golden_args()
    args.model = 'seq2seq'
    out = generate_gestures(args, StubSeq2Seq(), StubVocab(), audio, words, vid=1, seed_seq=seed, fade_out=fade)
    assert out.shape == g[tag].shape
    assert np.abs(out - g[tag]).max() < 2e-6, np.abs(out - g[tag]).max()


def test_batched_chains_equal_single_chains():
    """Clips of different lengths advanced in lock-step give the same result as one clip at a time."""
    from synthesize import generate_gestures, generate_gestures_batch
    args = golden_args()
    clips = []
    for i, sec in enumerate((9.3, 1.7, 6.1, 4.0)):
        audio, words, seed = make_clip(sec, seed=10 + i)
        clips.append(dict(audio=audio, words=words, vid=3 + i, seed_seq=seed))
    batch = generate_gestures_batch(args, StubGenerator(), StubVocab(), clips)
    for c, b in zip(clips, batch):
        single = generate_gestures(args, StubGenerator(), StubVocab(), c['audio'], c['words'], vid=c['vid'], seed_seq=c['seed_seq'])
        assert single.shape == b.shape and np.abs(single - b).max() < 1e-6


@pytest.mark.gpu
def test_driver_runs_real_generator_on_gpu():
    """The real PoseGenerator behind the driver: 3 clips of different lengths in one lock-step batch, finite poses, no per-window sync."""
    from gpu_util import build_ours
    from oracle import trimodal_oracle as O
    from synthesize import generate_gestures_batch
    dev = torch.device('cuda:0')
    cfg = O.HotPathConfig(n_words=400, n_speakers=16)
    args, G, D, _, _ = build_ours(cfg, dev)
    args.model, args.motion_resampling_framerate = 'multimodal_context', 15
    args.mean_dir_vec = golden_args().mean_dir_vec
    G.eval()
    clips = []
    for i, sec in enumerate((7.0, 2.0, 4.5)):
        audio, words, seed = make_clip(sec, seed=30 + i)
        clips.append(dict(audio=audio, words=words, vid=1 + i, seed_seq=seed))
    outs = generate_gestures_batch(args, G, StubVocab(), clips, fade_out=True)
    for o in outs:
        assert o.ndim == 2 and o.shape[1] == 27 and np.isfinite(o).all()
    assert outs[0].shape[0] > outs[2].shape[0] > outs[1].shape[0]
