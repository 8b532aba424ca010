"""Generates tests/golden/synthesize_driver.npz by executing the reference's OWN generate_gestures (scripts/synthesize.py:36-209,
extracted with `ast`: the module imports librosa / gentle / lmdb and cannot be imported) and DataPreprocessor.get_words_in_time_range
(scripts/data_loader/data_preprocessor.py:173-188) around a deterministic stub generator.  The stub (oracle.synthesize_stub) depends
on the seed poses, the audio slice, the word-index row and the speaker id, so the window schedule, the audio / text slicing, the seed
hand-off, the cross-fade and the fade-out are all pinned.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_synthesize"""
import argparse
import ast
import contextlib
import io
import math
import os
import random
import time

import numpy as np
import torch

from .make_golden import OUT
from .make_golden_eval import MEAN_DIR_VEC
from .synthesize_stub import StubGenerator, StubSeq2Seq, StubVocab, make_clip

SYN = '/root/reference/scripts/synthesize.py'
DPP = '/root/reference/scripts/data_loader/data_preprocessor.py'


def reference_generate_gestures():
    tree = ast.parse(open(DPP).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'DataPreprocessor'][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'get_words_in_time_range'][0]
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), DPP, 'exec'), ns)          # the @staticmethod decorator makes a staticmethod object

    class DataPreprocessor:
        get_words_in_time_range = ns['get_words_in_time_range']
    tree = ast.parse(open(SYN).read())
    gg = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'generate_gestures'][0]
    ns2 = {'torch': torch, 'np': np, 'math': math, 'random': random, 'time': time, 'device': torch.device('cpu'),
           'DataPreprocessor': DataPreprocessor, 'extract_melspectrogram': None}
    exec(compile(ast.Module(body=[gg], type_ignores=[]), SYN, 'exec'), ns2)
    return ns2['generate_gestures']


def golden_args():
    return argparse.Namespace(n_poses=34, n_pre_poses=4, motion_resampling_framerate=15, mean_dir_vec=MEAN_DIR_VEC, model='multimodal_context',
                              z_type='speaker')


def main():
    gg = reference_generate_gestures()
    args = golden_args()
    store = {}
    for tag, seconds, fade in (('short', 1.7, False), ('long', 9.3, False), ('fade', 6.1, True)):
        audio, words, seed = make_clip(seconds, seed=len(tag))
        with contextlib.redirect_stdout(io.StringIO()):
            out = gg(args, StubGenerator(), StubVocab(), audio, words, vid=7, seed_seq=seed, fade_out=fade)
        store[tag] = np.asarray(out)
        print(tag, np.asarray(out).shape)
    # the seq2seq baseline takes the same driver with its own call form and an extra cubic smoothing pass over every window boundary (:163-185)
    args.model = 'seq2seq'
    for tag, seconds, fade in (('s2s_short', 1.7, False), ('s2s_long', 9.3, False), ('s2s_fade', 6.1, True)):
        audio, words, seed = make_clip(seconds, seed=len(tag))
        with contextlib.redirect_stdout(io.StringIO()):
            out = gg(args, StubSeq2Seq(), StubVocab(), audio, words, vid=1, seed_seq=seed, fade_out=fade)
        store[tag] = np.asarray(out)
        print(tag, np.asarray(out).shape)
    np.savez(os.path.join(OUT, 'synthesize_driver.npz'), **store)


if __name__ == '__main__':
    main()
