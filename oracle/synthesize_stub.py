"""Deterministic stand-ins used to pin the long-form inference DRIVER (window schedule, slicing, hand-off, cross-fade) independently of
the generator's arithmetic.  TEST INFRASTRUCTURE ONLY.  Works on CPU tensors (reference execution) and CUDA tensors (our driver)."""
import numpy as np
import torch


class StubVocab:
    SOS_token, EOS_token, UNK_token = 1, 2, 3

    def get_word_index(self, word):
        return 4 + (sum(ord(c) for c in word) % 97)


class _Z:
    n_words = 11


class StubGenerator:
    """out[b,t,d] = smooth pattern + dependence on the seed poses, the audio slice, the word ids and the speaker id."""
    z_obj = _Z()

    def __call__(self, pre_seq, in_text_padded, in_audio, vid):
        B, T, D1 = pre_seq.shape
        D = D1 - 1
        dev = pre_seq.device
        t = torch.arange(T, device=dev, dtype=torch.float32).view(1, T, 1)
        d = torch.arange(D, device=dev, dtype=torch.float32).view(1, 1, D)
        base = 0.1 * torch.sin(0.37 * t + 0.5 * d)
        seed = 0.5 * pre_seq[:, :4, :-1].mean(dim=1, keepdim=True) * pre_seq[:, :1, -1:].clamp(0, 1)
        L = in_audio.shape[1] // T
        a = in_audio[:, :L * T].reshape(B, T, L).double().sum(dim=2).float().unsqueeze(2) * 1e-2
        w = in_text_padded.float().unsqueeze(2) * 1e-3
        v = (vid.float().view(B, 1, 1) * 0.01) if vid is not None else 0.0
        return base + seed + a + w + v, None, None, None

    def parameters(self):
        return iter([torch.zeros(1)])


class StubSeq2Seq:
    """seq2seq call form (synthesize.py:134-136): (in_text [1,L] ragged, lengths, seed poses [1,>=4,D], None) -> [1,T,D].  The per-window
    offset (word ids, sentence length) makes consecutive windows disagree at their boundary, which is what the seq2seq-only cubic
    smoothing of synthesize.py:163-185 acts on."""
    n_frames, pose_dim = 34, 27

    def __call__(self, in_text, lengths, poses, _):
        dev = in_text.device
        T, D = self.n_frames, self.pose_dim
        t = torch.arange(T, device=dev, dtype=torch.float32).view(1, T, 1)
        d = torch.arange(D, device=dev, dtype=torch.float32).view(1, 1, D)
        base = 0.1 * torch.sin(0.29 * t + 0.4 * d) + 0.002 * t * torch.cos(0.9 * d)
        seed = 0.5 * poses[:, :4].float().mean(dim=1, keepdim=True)
        w = (in_text.float().sum() % 13.0) * 0.01 + 0.003 * float(in_text.shape[1])
        return base + seed + w

    def parameters(self):
        return iter([torch.zeros(1)])


def make_clip(seconds, seed=0, sr=16000):
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    n = int(seconds * sr)
    audio = (0.1 * rng.standard_normal(n)).astype(np.float32)
    words, t = [], 0.05
    while t < seconds - 0.3:
        dur = float(rng.uniform(0.15, 0.5))
        words.append(['w%d' % int(rng.integers(0, 1000)), t, t + dur])
        t += dur + float(rng.uniform(0.0, 0.4))
    seed_seq = (0.1 * rng.standard_normal((4, 27))).astype(np.float32)
    return audio, words, seed_seq
