"""CPU: the C-ABI library loads and exports every symbol include/tg_b200.h declares; the drop-in nn.Modules carry the
reference's state_dict keys and shapes; host-side bookkeeping (flat arena ordering) is sound.  No kernel is launched."""
import ctypes
import os

import pytest
import torch

from oracle import synth
from oracle.make_golden import golden_cfg


def test_library_exports_every_header_symbol():
    from tgb200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.tg_version() >= 100
    sizes = (ctypes.c_int * 2)()
    loaded.tg_struct_sizes(sizes)
    assert sizes[0] == ctypes.sizeof(_lib.ConvGemm) and sizes[1] == ctypes.sizeof(_lib.ConvWgrad)


def test_header_has_no_torch_types():
    from tgb200 import _lib
    import re
    src = open(_lib.HEADER).read()
    code = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)          # signatures only, comments stripped
    assert 'torch' not in code.lower() and 'at::' not in code and 'Tensor' not in code and 'extern "C"' in code


def test_modules_have_reference_state_dict_keys():
    from gpu_util import build_ours
    cfg = golden_cfg()
    args, G, D, gsd, dsd = build_ours(cfg, None)            # strict=True load inside
    assert set(G.state_dict().keys()) == set(synth.with_tcn_aliases(gsd).keys())
    assert list(D.state_dict().keys()) == list(dsd.keys())
    from model.embedding_net import EmbeddingNet
    from gpu_util import make_args
    E = EmbeddingNet(make_args(cfg), cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None, 'pose')
    E.load_state_dict(synth.embedding_net_state_dict(cfg), strict=True)
    assert G.z_obj is not None and G.pre_length == 4 and G.gen_length == 30 and G.in_size == 108 and G.hidden_size == 300


def test_arena_order_puts_gru_directions_adjacent():
    from gpu_util import build_ours
    from tgb200.arena import ParamArena
    from tgb200.engine import gru_arena_order
    cfg = golden_cfg()
    args, G, D, _, _ = build_ours(cfg, None)
    for m in (G, D):
        names = [n for n, _ in m.named_parameters()]
        a = ParamArena(m, gru_arena_order(names))
        for l in range(4):
            assert a.adjacent(f'gru.weight_ih_l{l}', f'gru.weight_ih_l{l}_reverse')
            assert a.adjacent(f'gru.bias_ih_l{l}', f'gru.bias_ih_l{l}_reverse')
        a.ensure(torch.device('cpu'))                      # flat storage is plain torch memory: works without a GPU
        p = dict(m.named_parameters())['gru.weight_ih_l1']
        assert p.data_ptr() == a.flat.data_ptr() + 4 * a.offsets['gru.weight_ih_l1']
        assert p.grad is not None and p.grad.data_ptr() == a.grad.data_ptr() + 4 * a.offsets['gru.weight_ih_l1']
    sd = G.state_dict()
    assert sd['gru.weight_ih_l1'].shape == (900, 600)


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gesture-generation-from-trimodal-context_b200')
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
                # ... nor the test-only NumPy restatement of the C ABI (tests/cabi_emulator.py), nor anything else under tests/
                assert 'cabi_emulator' not in src and 'import tests' not in src and 'from tests' not in src, f


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from tgb200 import _lib
    from train_eval.train_gan import train_iter_gan
    with pytest.raises(_lib.TgError):
        train_iter_gan(None, 0, None, None, torch.zeros(1, 34, 27), None, None, None, None, None)
    # the embedding-model step functions and modules refuse CPU tensors as well
    import train_feature_extractor as tfx
    from model.embedding_net import EmbeddingNet
    from train_eval.train_joint_embed import eval_embed, train_iter_embed
    net = EmbeddingNet(None, 27, 34, None, None, None, 'pose').train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    x = torch.zeros(2, 34, 27)
    for call in (lambda: tfx.train_iter(None, 0, x, net, opt), lambda: train_iter_embed(None, 0, None, None, x, net, opt),
                 lambda: eval_embed(None, None, None, x, net), lambda: net(None, None, None, x, None)):
        with pytest.raises(_lib.TgError):
            call()


def test_seq2seq_module_keys_and_no_cpu_fallback():
    """Seq2SeqNet carries the reference's state_dict keys (strict load of the synthetic reference-keyed weights) and refuses CPU tensors."""
    import argparse
    from oracle.make_golden_seq2seq import golden_cfg as s2s_cfg
    from model.seq2seq_net import Seq2SeqNet
    from tgb200 import _lib
    from train_eval.train_seq2seq import train_iter_seq2seq
    cfg = s2s_cfg()
    args = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.1, n_pre_poses=cfg.n_pre_poses, GAN_noise_size=0,
                              loss_regression_weight=250.0, loss_kld_weight=0.1, loss_reg_weight=25.0)
    net = Seq2SeqNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None)
    sd = synth.seq2seq_state_dict(cfg)
    net.load_state_dict(sd, strict=True)
    assert set(net.state_dict().keys()) == set(sd.keys())
    if not torch.cuda.is_available():
        inp = synth.seq2seq_inputs(cfg, 4, seed=1, max_len=6)
        with pytest.raises(_lib.TgError):
            net(inp['in_text'], inp['lengths'], inp['target'], None)
        with pytest.raises(_lib.TgError):
            train_iter_seq2seq(args, 0, inp['in_text'], inp['lengths'], inp['target'], net.train(), torch.optim.Adam(net.parameters()))


def test_seq2seq_launch_plan_trace():
    """Host logic of the seq2seq plan without a GPU: TGB200_TRACE_ONLY records which C-ABI entry points one training iteration would call."""
    import subprocess
    import sys
    code = r'''
import sys, argparse, collections
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
from oracle import synth
from oracle.make_golden_seq2seq import golden_cfg
from model.seq2seq_net import Seq2SeqNet
from train_eval.train_seq2seq import train_iter_seq2seq
from tgb200 import _lib
cfg = golden_cfg()
args = argparse.Namespace(hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=0.1, n_pre_poses=cfg.n_pre_poses, GAN_noise_size=0,
                          loss_regression_weight=250.0, loss_kld_weight=0.1, loss_reg_weight=25.0)
net = Seq2SeqNet(args, cfg.pose_dim, cfg.n_poses, cfg.n_words, cfg.wordembed_dim, None).train()
inp = synth.seq2seq_inputs(cfg, 6, seed=3, max_len=9)
train_iter_seq2seq(args, 0, inp['in_text'], inp['lengths'], inp['target'], net, torch.optim.Adam(net.parameters(), lr=1e-4))
c = collections.Counter(_lib.trace)
T, Tm, L = cfg.n_poses, 9, cfg.n_layers
assert c['tg_attn_fwd'] == T - 1 and c['tg_attn_bwd'] == T - 1, c
assert c['tg_gru_gates_fwd'] == (T - 1) * L + Tm * 2 * L and c['tg_gru_gates_bwd'] == c['tg_gru_gates_fwd'], c
assert c['tg_s2s_loss'] == 1 and c['tg_sumsq_f64'] == 1 and c['tg_clip_scale'] == 1 and c['tg_adam_flat'] == 1, c
assert c['tg_bn_finalize'] == T - 1 and c['tg_bn_bwd_apply'] == T - 1, c
print('ok', sum(c.values()))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TGB200_TRACE_ONLY='1')
    r = subprocess.run([sys.executable, '-c', code % (root, os.path.join(root, 'gesture-generation-from-trimodal-context_b200'))], env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]


def test_checkpoint_layout_round_trip(tmp_path):
    """the reference checkpoint {'args','epoch','lang_model','speaker_model','pose_dim','gen_dict','dis_dict'} (train.py:153-157)
    written from our modules loads the way utils/train_utils.py:166-183 loads it: unpickle (model.vocab.Vocab resolves to our class),
    train.init_model(...), load_state_dict(strict)."""
    import argparse
    import train
    from model import vocab
    cfg = golden_cfg()
    lang = vocab.Vocab('words')
    for i in range(cfg.n_words - lang.n_words):
        lang.index_word('w%d' % i)
    lang.word_embedding_weights = None
    spk = vocab.Vocab('vid', insert_default_tokens=False)
    for i in range(cfg.n_speakers - 1):
        spk.index_word('spk%d' % i)
    args = argparse.Namespace(model='multimodal_context', n_poses=cfg.n_poses, n_pre_poses=cfg.n_pre_poses, wordembed_dim=cfg.wordembed_dim,
                              hidden_size=cfg.hidden_size, n_layers=cfg.n_layers, dropout_prob=cfg.dropout_prob, freeze_wordembed=False,
                              z_type='speaker', input_context='both')
    G, D, _ = train.init_model(args, lang, spk, cfg.pose_dim, torch.device('cpu'))
    path = str(tmp_path / 'ckpt.bin')
    torch.save({'args': args, 'epoch': 7, 'lang_model': lang, 'speaker_model': spk, 'pose_dim': cfg.pose_dim,
                'gen_dict': G.state_dict(), 'dis_dict': D.state_dict()}, path)
    ck = torch.load(path, map_location='cpu', weights_only=False)
    assert set(ck) == {'args', 'epoch', 'lang_model', 'speaker_model', 'pose_dim', 'gen_dict', 'dis_dict'} and ck['epoch'] == 7
    assert isinstance(ck['lang_model'], vocab.Vocab) and ck['speaker_model'].n_words == spk.n_words
    G2, D2, _ = train.init_model(ck['args'], ck['lang_model'], ck['speaker_model'], ck['pose_dim'], torch.device('cpu'))
    G2.load_state_dict(ck['gen_dict'])                       # strict=True by default, train_utils.py:178
    D2.load_state_dict(ck['dis_dict'])
    assert all(torch.equal(a, b) for a, b in zip(G.state_dict().values(), G2.state_dict().values()))
    assert G2.z_obj is ck['speaker_model']                   # what utils.train_utils.get_speaker_model(generator) returns (:152-164)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle on the host cores) prints ONE JSON line with the contract's keys; no GPU involved."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '8'],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == 'G+D train samples/s' and j['unit'] == 'samples/s' and j['higher_is_better'] is True
    assert j['value'] > 0 and j['steps'] == 1 and j['n_gpus'] == 1 and 'workload' in j['config']
    assert j['cpu_baseline']['kind'] == 'port' and j['cpu_baseline']['cores'] >= 1 and j['cpu_baseline']['value'] == j['value']
    assert j['e2e'] == {'value': j['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
