"""CPU (only where /root/reference is present, i.e. in the build container): the UNMODIFIED reference step function
scripts/train_eval/train_gan.py::train_iter_gan - torch autograd, loss.backward(), the caller's torch.optim.Adam - driven over OUR
PoseGenerator / ConvDiscriminator through their nn.Module API (launch plans on the NumPy C-ABI emulator), compared with the golden the same
function produced on the reference's own modules (tests/golden/train_e11.npz, train_e0.npz).  This is the drop-in claim of SURVEY 8b taken
literally: the reference code runs on our modules and computes what it computed on its own."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import cabi_emulator
import test_gpu_parity as GP
from conftest import GOLDEN
from gpu_util import build_ours, masks_to_ours
from oracle import synth
from oracle.make_golden import digest, golden_cfg

REF = '/root/reference/scripts/train_eval/train_gan.py'
CPU = torch.device('cpu')


@pytest.mark.skipif(not os.path.exists(REF), reason='the reference tree is only present in the build container')
@pytest.mark.parametrize('tag,epoch,use_masks', [('train_e11', 11, True), ('train_e0', 0, False)])
def test_unmodified_reference_train_iter_gan_runs_on_our_modules(tag, epoch, use_masks):
    from tgb200 import config
    spec = importlib.util.spec_from_file_location('ref_train_gan', REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    orig_randperm = torch.randperm
    try:
        with cabi_emulator.installed():
            cfg = golden_cfg()
            g = np.load(os.path.join(GOLDEN, tag + '.npz'))
            args, G, D, _, _ = build_ours(cfg, CPU)
            G.train(); D.train()
            inp = synth.make_inputs(cfg, 3, seed=1)
            noise = synth.golden_noise(cfg, 3, 2, use_masks)
            fwd_ids = [0, 1, 2] if epoch > cfg.loss_warmup else [1, 2]
            g_queue = [(noise.eps[i], masks_to_ours(noise.g_masks[i], CPU) if noise.g_masks[i] else {}) for i in fwd_ids]
            g_forward, d_forward = G.forward, D.forward

            def g_fwd(*a, **k):                                  # the reference calls the generator two or three times per step
                eps, masks = g_queue.pop(0)
                G.set_noise(eps=eps, masks=masks)
                return g_forward(*a, **k)

            def d_fwd(*a, **k):
                D.set_noise(masks={})                            # golden: GRU inter-layer dropout switched off (it cannot take a mask)
                return d_forward(*a, **k)
            G.forward, D.forward = g_fwd, d_fwd
            torch.randperm = lambda n, *a, **k: noise.perm.clone()
            g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
            d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
            ret = ref.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
            assert not g_queue
            assert set('loss_' + k for k in ret) == set(k for k in g.files if k.startswith('loss_'))
            for k, v in ret.items():
                r = float(g['loss_' + k])
                assert abs(v - r) <= GP.TOL * abs(r) + 1e-6, (k, v, r)
            for k, p in G.named_parameters():
                GP._digest_close(digest(p.grad), g['ggrad/' + k], 3e-4)
            gsd, dsd = G.state_dict(), D.state_dict()
            for k in g.files:
                if k.startswith('gpost/'):
                    GP._post_close(digest(gsd[k[6:]]), g[k], cfg.learning_rate, k[6:] in GP.ZERO_GRAD_KEYS)
                if k.startswith('dpost/'):
                    GP._post_close(digest(dsd[k[6:]]), g[k], cfg.learning_rate * cfg.discriminator_lr_weight, k[6:] in GP.ZERO_GRAD_KEYS)
    finally:
        torch.randperm = orig_randperm
        config.set_mode(old_mode); config.set_graphs(old_graphs)


FX = '/root/reference/scripts/train_feature_extractor.py'


@pytest.mark.skipif(not os.path.exists(FX), reason='the reference tree is only present in the build container')
def test_unmodified_reference_autoencoder_train_iter_runs_on_our_module():
    """scripts/train_feature_extractor.py::train_iter and train_eval/train_joint_embed.py::train_iter_embed, unmodified (autograd +
    torch Adam), over OUR EmbeddingNet(mode='pose'): two consecutive steps reproduce the golden the reference produced on its own module."""
    import ae_checks
    from oracle.make_golden_ae import reference_train_iter
    from tgb200 import config
    train_iter = reference_train_iter()                      # extracted from the reference source with ast, executed as is
    spec = importlib.util.spec_from_file_location('ref_joint', '/root/reference/scripts/train_eval/train_joint_embed.py')
    ref_joint = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_joint)
    old_graphs = config.set_graphs(False)
    try:
        with cabi_emulator.installed():
            g = np.load(os.path.join(GOLDEN, 'ae_train.npz'))
            cfg, net, opt = ae_checks.build(CPU)
            net.train()
            for step in (1, 2):
                ret = train_iter(None, 0, torch.from_numpy(g[f'fx{step}/target']), net, opt)
                ae_checks.check_against_golden(g, f'fx{step}', net, ret['loss'], step, float(g['lr']), 2e-5)
            cfg, net, opt = ae_checks.build(CPU)
            net.train()
            import argparse
            ret = ref_joint.train_iter_embed(argparse.Namespace(n_pre_poses=cfg.n_pre_poses), 0, None, None, torch.from_numpy(g['fx1/target']), net, opt)
            ae_checks.check_against_golden(g, 'je', net, ret['loss'], 1, float(g['lr']), 2e-5)
    finally:
        config.set_graphs(old_graphs)


S2S = '/root/reference/scripts/train_eval/train_seq2seq.py'


@pytest.mark.skipif(not os.path.exists(S2S), reason='the reference tree is only present in the build container')
def test_unmodified_reference_train_iter_seq2seq_runs_on_our_module():
    """scripts/train_eval/train_seq2seq.py::train_iter_seq2seq, unmodified (custom_loss in torch, loss.backward(), clip_grad_norm_,
    torch Adam), over OUR Seq2SeqNet: two consecutive steps reproduce the golden the reference produced on its own module."""
    import test_gpu_seq2seq as GS
    from oracle import seq2seq_oracle as S
    from tgb200 import config
    spec = importlib.util.spec_from_file_location('ref_train_seq2seq', S2S)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed():
            cfg = GS.golden_cfg()
            g = np.load(os.path.join(GOLDEN, 'seq2seq_step.npz'))
            args, net = GS._build(cfg, CPU)
            inp = synth.seq2seq_inputs(cfg, 6, seed=3, max_len=9)
            net.train()
            optim = torch.optim.Adam(net.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
            for it in range(2):
                ret = ref.train_iter_seq2seq(args, 0, inp['in_text'], inp['lengths'], inp['target'], net, optim)
                r = float(g[f'loss{it}'])
                assert abs(ret['loss'] - r) <= GS.FP32_TOL * abs(r), (it, ret['loss'], r)
                for k, p in net.named_parameters():
                    GS._digest_close(digest(p.grad), g[f'grad{it}/' + k], 1e-3)
                for k, v in net.state_dict().items():
                    rr = g[f'post{it}/' + k]
                    if k.endswith('num_batches_tracked'):
                        assert int(v) == int(rr[1])
                    elif S.is_param(k):
                        GS._post_close(digest(v), rr, cfg.learning_rate, noisy=(it > 0 or k == 'decoder.decoder.pre_linear.0.bias'))
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


@pytest.mark.skipif(not os.path.exists(FX), reason='the reference tree is only present in the build container')
def test_unmodified_reference_train_iter_embed_runs_on_our_joint_embedding_module():
    """train_joint_embed.py::train_iter_embed(mode='random'), unmodified, over OUR EmbeddingNet(mode='random') with the coin patched to
    'speech' then 'pose' (as the golden): losses, gradients, post-Adam weights and - the subtle part - torch.optim.Adam leaving the
    parameters of the branch that was not decoded alone."""
    import random
    import joint_checks
    from oracle.make_golden import digest as dg
    from test_oracle_ae_golden import digest_close, post_close
    from test_oracle_joint_golden import is_zero_grad
    from tgb200 import config
    spec = importlib.util.spec_from_file_location('ref_joint2', '/root/reference/scripts/train_eval/train_joint_embed.py')
    ref_joint = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_joint)
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    coins = [0.9, 0.1]
    orig = random.random
    random.random = lambda: coins.pop(0)
    try:
        with cabi_emulator.installed():
            g = np.load(os.path.join(GOLDEN, 'joint_embed.npz'))
            cfg, args, net, opt = joint_checks.build(CPU)
            net.train()
            for step in (1, 2):
                noise = synth.golden_noise(cfg, joint_checks.B, 70 + step, True)
                e = noise.eps[0].repeat(1, 2)[:, :32].contiguous()
                net.joint_engine().noise = dict(eps=e, masks=masks_to_ours(noise.g_masks[0], CPU), gru_masks=[None] * 4)
                data = synth.make_inputs(cfg, joint_checks.B, seed=63 + step)
                ret = ref_joint.train_iter_embed(args, 0, data['in_text'], data['in_audio'], data['target'], net, opt, mode='random')
                tag = f'step{step}'
                r = float(g[f'{tag}/loss'])
                assert abs(ret['loss'] - r) <= (2e-5 if step == 1 else 4e-4) * abs(r), (tag, ret['loss'], r)
                for k, p in net.named_parameters():
                    assert (p.grad is not None) == bool(g[f'{tag}/hasgrad/{k}']) or k.startswith('pose_encoder.fc_logvar'), (tag, k)
                    if p.grad is None or is_zero_grad(k) or not bool(g[f'{tag}/hasgrad/{k}']):
                        continue
                    digest_close(dg(p.grad), g[f'{tag}/grad/{k}'], 1e-4 if step == 1 else 5e-3)
                for k, v in net.state_dict().items():
                    if ('.tcn.network.' in k and ('.net.0.' in k or '.net.4.' in k)) or is_zero_grad(k) or 'running' in k or k.endswith('num_batches_tracked'):
                        continue
                    post_close(dg(v), g[f'{tag}/post/{k}'], float(g['lr']), step)
        assert not coins
    finally:
        random.random = orig
        config.set_mode(old_mode); config.set_graphs(old_graphs)


@pytest.mark.skipif(not os.path.exists('/root/reference/scripts/synthesize.py'), reason='the reference tree is only present in the build container')
def test_unmodified_reference_generate_gestures_drives_our_generator():
    """scripts/synthesize.py::generate_gestures (extracted with ast, executed as is: one forward per window, per-window .cpu()) calling OUR
    PoseGenerator at the synthesize.py:131 call site gives what OUR long-form driver gives for the same clip: both draw the same Philox
    noise sequence (one eval forward per window on a fresh, identically seeded module)."""
    import contextlib
    import io
    from oracle import trimodal_oracle as O
    from oracle.make_golden_synthesize import golden_args, reference_generate_gestures
    from oracle.synthesize_stub import StubVocab, make_clip
    from synthesize import generate_gestures
    from tgb200 import config
    gg = reference_generate_gestures()
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed():
            cfg = O.HotPathConfig(n_words=300, n_speakers=12)
            audio, words, seed = make_clip(4.4, seed=77)                         # two windows
            outs = []
            for driver in (gg, generate_gestures):
                torch.manual_seed(11)                                             # module noise seed
                args, G, D, _, _ = build_ours(cfg, CPU)
                a = golden_args()
                for k in ('n_poses', 'n_pre_poses', 'motion_resampling_framerate', 'mean_dir_vec', 'model', 'z_type'):
                    setattr(args, k, getattr(a, k))
                G.eval()
                with contextlib.redirect_stdout(io.StringIO()):
                    outs.append(np.asarray(driver(args, G, StubVocab(), audio, words, vid=5, seed_seq=seed, fade_out=True)))
            assert outs[0].shape == outs[1].shape and outs[0].shape[0] > 34 and np.isfinite(outs[0]).all()
            assert np.abs(outs[0] - outs[1]).max() < 2e-6, np.abs(outs[0] - outs[1]).max()
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


ESE = '/root/reference/scripts/model/embedding_space_evaluator.py'


@pytest.mark.skipif(not os.path.exists(ESE), reason='the reference tree is only present in the build container')
def test_unmodified_reference_embedding_space_evaluator_uses_our_embedding_net(tmp_path):
    """scripts/model/embedding_space_evaluator.py, unmodified (its `from model.embedding_net import EmbeddingNet` resolves to OUR module):
    constructor from a checkpoint file (:16-28), push_samples (:45-61), get_scores (:74-101) -> the golden FGD of the reference run."""
    import argparse
    import sys
    import types
    from scipy import linalg as sl
    from tgb200 import config
    sys.modules.setdefault('umap', types.ModuleType('umap'))                      # imported at module top, only used for visualisation
    spec = importlib.util.spec_from_file_location('ref_ese', ESE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    import model.embedding_net as ours
    assert ref.EmbeddingNet is ours.EmbeddingNet

    class _L:                                                                     # SciPy >= 1.18 dropped disp= (SURVEY 8c shim 3)
        @staticmethod
        def sqrtm(a, disp=True):
            r = sl.sqrtm(a)
            return r if disp else (r, 0.0)
    ref.linalg = _L
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'embedding_fgd.npz'))
    path = str(tmp_path / 'embed_net.bin')
    torch.save({'pose_dim': cfg.pose_dim, 'gen_dict': synth.embedding_net_state_dict(cfg)}, path)
    args = argparse.Namespace(n_pre_poses=cfg.n_pre_poses, n_poses=cfg.n_poses, wordembed_dim=cfg.wordembed_dim, hidden_size=cfg.hidden_size,
                              n_layers=cfg.n_layers, dropout_prob=0.3, freeze_wordembed=False)
    lang = argparse.Namespace(n_words=cfg.n_words, word_embedding_weights=None)
    rng = np.random.Generator(np.random.PCG64(77))
    real = torch.from_numpy((0.5 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim))).astype(np.float32))
    fake = torch.from_numpy((1.0 * rng.standard_normal((256, cfg.n_poses, cfg.pose_dim)) + 0.3).astype(np.float32))
    old_graphs = config.set_graphs(False)
    try:
        with cabi_emulator.installed(), torch.no_grad():
            ev = ref.EmbeddingSpaceEvaluator(args, path, lang, CPU)
            for i in range(0, 256, 64):
                ev.push_samples(None, None, fake[i:i + 64], real[i:i + 64])
            fgd, feat_dist = ev.get_scores()
    finally:
        config.set_graphs(old_graphs)
    assert ev.get_no_of_samples() == 4
    assert abs(fgd - float(g['fgd'])) <= 1e-3 * abs(float(g['fgd'])), (fgd, float(g['fgd']))
    assert abs(feat_dist - float(g['feat_dist'])) <= 1e-4 * abs(float(g['feat_dist']))


@pytest.mark.skipif(not os.path.exists(REF), reason='the reference tree is only present in the build container')
@pytest.mark.parametrize('gan_w,reg_w,epoch', [(0.0, 0.05, 11), (5.0, 0.0, 11), (0.0, 0.0, 11), (0.0, 0.0, 0)])
def test_native_step_equals_reference_step_for_other_loss_weight_settings(gan_w, reg_w, epoch):
    """train_gan.py:27,58,88 branch on loss_gan_weight / loss_reg_weight / the warm-up: our native train_iter_gan and the UNMODIFIED
    reference train_iter_gan (autograd over our modules) must agree on which losses are returned, their values and every gradient."""
    from conftest import rel_l2
    from tgb200 import config
    from train_eval import train_gan as TG
    spec = importlib.util.spec_from_file_location('ref_train_gan2', REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    orig_randperm = torch.randperm
    try:
        with cabi_emulator.installed():
            cfg = golden_cfg()
            inp = synth.make_inputs(cfg, 3, seed=1)
            noise = synth.golden_noise(cfg, 3, 2, False)
            results = []
            for native in (True, False):
                args, G, D, _, _ = build_ours(cfg, CPU)
                args.loss_gan_weight, args.loss_reg_weight = gan_w, reg_w
                G.train(); D.train()
                g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
                d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
                after, do_d, do_div, _ = TG._flags(args, epoch)
                fwd_ids = ([0] if do_d else []) + [1] + ([2] if do_div else [])
                if native:
                    TG.inject_noise(TG.StepNoise(eps=[noise.eps[i] for i in fwd_ids], perm=noise.perm, g_masks=[{}] * len(fwd_ids), d_masks=[{}, {}, {}]))
                    ret = TG.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
                else:
                    queue = [noise.eps[i] for i in fwd_ids]
                    g_forward, d_forward = G.forward, D.forward

                    def g_fwd(*a, **k):
                        G.set_noise(eps=queue.pop(0), masks={})
                        return g_forward(*a, **k)

                    def d_fwd(*a, **k):
                        D.set_noise(masks={})
                        return d_forward(*a, **k)
                    G.forward, D.forward = g_fwd, d_fwd
                    torch.randperm = lambda n, *a, **k: noise.perm.clone()
                    ret = ref.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
                    assert not queue
                results.append((ret, {k: p.grad.clone() for k, p in G.named_parameters()}))
            (a, ga), (b, gb) = results
            assert set(a) == set(b), (a, b)
            for k in a:
                assert abs(a[k] - b[k]) <= 1e-5 * abs(b[k]) + 1e-7, (k, a[k], b[k])
            for k in ga:
                if gb[k].norm() > 1e-6 and k not in GP.ZERO_GRAD_KEYS:
                    assert rel_l2(ga[k], gb[k]) < 1e-4, k
    finally:
        torch.randperm = orig_randperm
        config.set_mode(old_mode); config.set_graphs(old_graphs)


TRAIN = '/root/reference/scripts/train.py'


@pytest.mark.skipif(not os.path.exists(TRAIN), reason='the reference tree is only present in the build container')
@pytest.mark.parametrize('model', ['multimodal_context', 'seq2seq'])
def test_unmodified_reference_evaluate_testset_agrees_with_ours(model):
    """scripts/train.py::evaluate_testset (extracted with ast together with the reference's AverageMeter, get_speaker_model and
    convert_dir_vec_to_pose; the module itself imports matplotlib / lmdb) run over OUR generator and OUR EmbeddingSpaceEvaluator, vs OUR
    device-side evaluate_testset on fresh, identically seeded modules: same loss, joint MAE, FGD and feature distance."""
    import argparse
    import ast
    import logging
    import random
    import time
    import torch.nn.functional as F
    import test_gpu_seq2seq as GS
    from model import vocab
    from model.embedding_net import EmbeddingNet
    from model.embedding_space_evaluator import EmbeddingSpaceEvaluator
    from oracle import seq2seq_oracle as S
    from oracle import trimodal_oracle as O
    from oracle.make_golden_eval import MEAN_DIR_VEC, reference_convert
    from tgb200 import config
    from train_eval.evaluate import evaluate_testset as ours
    from train_eval.train_joint_embed import eval_embed

    def extract(path, names, ns):
        tree = ast.parse(open(path).read())
        keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
        assert len(keep) == len(names), names
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
        return ns
    um = extract('/root/reference/scripts/utils/train_utils.py', ['get_speaker_model'], {'vocab': vocab})
    am = extract('/root/reference/scripts/utils/average_meter.py', ['AverageMeter'], {})
    utils_ns = argparse.Namespace(train_utils=argparse.Namespace(get_speaker_model=um['get_speaker_model']))
    ns = extract(TRAIN, ['evaluate_testset'], {'torch': torch, 'np': np, 'F': F, 'time': time, 'random': random, 'logging': logging, 'device': CPU,
                                                'AverageMeter': am['AverageMeter'], 'utils': utils_ns, 'eval_embed': eval_embed,
                                                'convert_dir_vec_to_pose': reference_convert()})
    ref_eval = ns['evaluate_testset']
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed():
            cfg = O.HotPathConfig(n_words=200, n_speakers=12)
            s_cfg = S.Seq2SeqConfig(n_words=200)
            batches = []
            for i, B in enumerate((5, 3)):                                       # unequal batch sizes: the averages are sample-weighted
                inp = synth.make_inputs(cfg, B, seed=40 + i)
                s2s = synth.seq2seq_inputs(s_cfg, B, seed=50 + i, max_len=7)
                batches.append((s2s['in_text'], s2s['lengths'], inp['in_text'], None, inp['target'], inp['in_audio'], torch.zeros(B, 2, 2), None))
            results = []
            for fn in (ref_eval, ours):
                torch.manual_seed(5); random.seed(5)
                if model == 'multimodal_context':
                    args, G, _, _, _ = build_ours(cfg, CPU)
                    loss_fn = None
                else:
                    args, G = GS._build(s_cfg, CPU)
                    loss_fn = torch.nn.L1Loss()
                args.model, args.n_poses, args.n_pre_poses, args.mean_dir_vec = model, cfg.n_poses, cfg.n_pre_poses, MEAN_DIR_VEC
                enet = EmbeddingNet(None, cfg.pose_dim, cfg.n_poses, None, None, None, 'pose')
                enet.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
                ev = EmbeddingSpaceEvaluator.from_net(enet, cfg.n_pre_poses, CPU)
                G.train()
                results.append(fn([tuple(t.clone() if torch.is_tensor(t) else t for t in b) for b in batches], G, loss_fn, ev, args))
                assert G.training
            a, b = results
            assert set(a) == set(b) == {'loss', 'joint_mae', 'frechet', 'feat_dist'}
            for k in a:
                assert abs(a[k] - b[k]) <= 2e-4 * abs(a[k]) + 1e-7, (k, a[k], b[k])
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)
