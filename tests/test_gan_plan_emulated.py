"""CPU: the HEADLINE launch plan - train_iter_gan in strict-fp32 mode (tgb200/engine.py, train_eval/train_gan.py: the 3*B generator
sweep, the discriminator passes, every hand-derived backward, both flat Adam steps) - executed on the NumPy restatement of the C-ABI
entries it calls (tests/cabi_emulator.py) and held to the same reference-executed goldens as on the GPU: the assertions are the GPU
tests' own (tests/test_gpu_parity.py), called here with CPU tensors.  This checks the HOST logic without a GPU; the kernels themselves
are checked on the B200."""
import pytest
import torch

import cabi_emulator
import test_gpu_parity as GP

CPU = torch.device('cpu')


@pytest.fixture()
def emu_fp32():
    from tgb200 import config
    old_mode, old_graphs = config.set_mode('fp32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed() as e:
            yield e
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


@pytest.mark.parametrize('tag,epoch,use_masks', [('train_e11', 11, True), ('train_e0', 0, False)])
def test_train_iter_gan_plan_vs_reference_golden(emu_fp32, tag, epoch, use_masks):
    GP.test_train_iter_vs_reference_golden(CPU, tag, epoch, use_masks)
    # the generator's Adam is two launches (the recurrent layers' range as soon as their gradients are final, the rest at the end), D's is one
    assert emu_fp32.calls.count('tg_adam_flat') == (3 if epoch > 10 else 2)
    assert 'tg_gru_layer_bwd' in emu_fp32.calls and 'tg_gen_losses' in emu_fp32.calls


def test_forward_eval_plan_vs_reference_golden(emu_fp32):
    GP.test_forward_eval_vs_reference_golden(CPU)


def test_train_iter_gan_all_dropout_masks_vs_fp64_oracle(emu_fp32):
    """Every dropout mask injected, incl. the GRU inter-layer masks the reference cannot take (hence the fp64 oracle); small batch."""
    GP.test_train_iter_full_size_vs_oracle(CPU, 4, 11, 2000, 50)


def test_module_api_autograd_plan(emu_fp32):
    """The torch.autograd glue of the nn.Module API (model/*.py): loss.backward() through our modules vs oracle autograd."""
    GP.test_module_api_autograd_matches_oracle(CPU)


@pytest.mark.parametrize('ctx,zm', [('audio', 'speaker'), ('text', 'random'), ('none', None)])
def test_constructor_variant_plans(emu_fp32, ctx, zm):
    GP.test_constructor_variants_vs_reference_golden(CPU, ctx, zm)


def test_embedding_net_and_fgd_plan(emu_fp32):
    GP.test_embedding_net_and_fgd_vs_reference_golden(CPU)


@pytest.fixture()
def emu_fast():
    from tgb200 import config
    old_mode, old_graphs = config.set_mode('tf32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed() as e:
            yield e
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


def test_fast_mode_plans(emu_fast):
    """The DEFAULT (tf32) mode routes the large GEMMs, weight gradients, the generator GRU and WavEncoder conv2-4 through the
    tensor-core entries (window views, two-tap causal GEMM, accumulating-tap data gradients, MN-major weight gradients): same goldens /
    oracle, other plan."""
    GP.test_fast_mode_forward_eval_vs_reference_golden(CPU)
    GP.test_fast_mode_train_iter_full_size_vs_oracle(CPU, 4, 11, 2000, 50)
    for sym in ('tg_gemm_tf32', 'tg_wgrad_tf32', 'tg_gru_layer_fwd_tf32', 'tg_gru_layer_bwd_tf32', 'tg_conv_dgrad_tf32', 'tg_conv1_wgrad'):
        assert sym in emu_fast.calls, sym


@pytest.mark.parametrize('B', [1, 5])
def test_small_and_odd_batches(emu_fast, B):
    """Edge cases of the launch plan: a single clip (every 'batch' statistic is over one clip's frames) and a batch that is not a multiple
    of any tile size, full G+D iteration with every dropout mask vs the fp64 oracle."""
    GP.test_train_iter_full_size_vs_oracle(CPU, B, 11, 2000, 50)


def test_fast_mode_error_budget_under_tf32_operand_truncation():
    """The emulator can cut the operands of the tensor-core entries to TF32 the way the hardware does when it is fed raw fp32 words
    (low 13 mantissa bits dropped).  With that model the default-mode plan must stay inside the north-star's 1e-2 budget against the fp64
    oracle (the test asserts it) - and lands where the B200 does: 2.6e-3 on the poses predicted here, 2.6e-3 measured."""
    from tgb200 import config
    old_mode, old_graphs = config.set_mode('tf32'), config.set_graphs(False)
    try:
        with cabi_emulator.installed(tf32_round='trunc'):
            GP.test_fast_mode_train_iter_full_size_vs_oracle(CPU, 8, 11, 2000, 50)       # asserts losses / poses <= 1e-2, worst gradient <= 5e-2
    finally:
        config.set_mode(old_mode); config.set_graphs(old_graphs)


@pytest.mark.parametrize('kw', [dict(n_layers=2), dict(hidden_size=128, wordembed_dim=128), dict(n_layers=3, hidden_size=64, wordembed_dim=64),
                                dict(n_pre_poses=2)], ids=lambda kw: ','.join('%s=%s' % it for it in kw.items()))
def test_other_hyper_parameters(emu_fast, kw):
    """The plan is not hard-wired to config/multimodal_context.yml: other layer counts, hidden / embedding widths (the generator GRU then
    takes the single-CTA or the tensor-core kernel family depending on H) and seed-pose counts, full G+D iteration vs the fp64 oracle."""
    from gpu_util import build_ours, masks_to_ours
    from conftest import rel_l2
    from oracle import synth
    from oracle import trimodal_oracle as O
    from train_eval import train_gan as TG
    cfg = O.HotPathConfig(n_words=300, n_speakers=12, **kw)
    B, epoch = 3, 11
    args, G, D, gsd, dsd = build_ours(cfg, CPU)
    G.train(); D.train()
    inp = synth.make_inputs(cfg, B, seed=3)
    noise = synth.make_noise(cfg, B, seed=4, dropout=True)
    ref = GP._oracle_step(cfg, epoch, gsd, dsd, inp, noise, CPU)
    g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
    TG.inject_noise(TG.StepNoise(eps=list(noise.eps), perm=noise.perm, g_masks=[masks_to_ours(m, CPU) for m in noise.g_masks],
                                 d_masks=[masks_to_ours(m, CPU) for m in noise.d_masks]))
    ret = TG.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
    assert set(ret) == set(ref['losses'])
    for k, v in ret.items():
        assert abs(v - ref['losses'][k]) <= 1e-5 * abs(ref['losses'][k]) + 1e-7, (k, v, ref['losses'][k])
    for k, p in G.named_parameters():
        r = ref['g_grads'][k]
        if r.norm() > 1e-6:
            assert rel_l2(p.grad, r) < 1e-4, k


def test_pretrained_frozen_word_embeddings(emu_fast):
    """args.freeze_wordembed=True with a pre-trained embedding matrix (multimodal_context_net.py:38-41, the fastText path): the table is not
    a trainable parameter - it stays out of the flat arena and of the optimiser, receives no scatter-add and does not move."""
    from gpu_util import make_args, masks_to_ours
    from conftest import rel_l2
    from model import vocab
    from model.multimodal_context_net import ConvDiscriminator, PoseGenerator
    from oracle import synth
    from oracle import trimodal_oracle as O
    from train_eval import train_gan as TG
    cfg = O.HotPathConfig(n_words=300, n_speakers=12)
    args = make_args(cfg)
    args.freeze_wordembed = True
    spk = vocab.Vocab('vid', insert_default_tokens=False)
    while spk.n_words < cfg.n_speakers:
        spk.index_word('s%d' % spk.n_words)
    gsd, dsd = synth.generator_state_dict(cfg), synth.discriminator_state_dict(cfg)
    emb = gsd['text_encoder.embedding.weight'].numpy().copy()
    G = PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, emb, z_obj=spk)
    D = ConvDiscriminator(cfg.pose_dim)
    G.load_state_dict(synth.with_tcn_aliases(gsd), strict=True); D.load_state_dict(dsd, strict=True)
    assert not G.text_encoder.embedding.weight.requires_grad
    G.train(); D.train()
    B = 3
    inp, noise = synth.make_inputs(cfg, B, seed=3), synth.make_noise(cfg, B, seed=4, dropout=True)
    ref = GP._oracle_step(cfg, 11, gsd, dsd, inp, noise, CPU)
    g_opt = torch.optim.Adam([p for p in G.parameters() if p.requires_grad], lr=cfg.learning_rate, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
    TG.inject_noise(TG.StepNoise(eps=list(noise.eps), perm=noise.perm, g_masks=[masks_to_ours(m, CPU) for m in noise.g_masks],
                                 d_masks=[masks_to_ours(m, CPU) for m in noise.d_masks]))
    ret = TG.train_iter_gan(args, 11, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
    for k, v in ret.items():
        assert abs(v - ref['losses'][k]) <= 1e-5 * abs(ref['losses'][k]) + 1e-7, k
    assert torch.equal(G.text_encoder.embedding.weight.data, torch.from_numpy(emb))
    assert emu_fast.calls.count('tg_embedding_scatter_add') == 1           # the speaker embedding only (2 with a trainable word table)
    for k, p in G.named_parameters():
        if p.requires_grad and ref['g_grads'][k].norm() > 1e-6:
            assert rel_l2(p.grad, ref['g_grads'][k]) < 1e-4, k


def test_changing_batch_size_between_steps(emu_fast):
    """Workspaces are cached per shape: a step at batch 3 AFTER a step at batch 5 (and an eval forward at batch 2 in between) must give the
    losses and gradients of the same step on a fresh copy of the model - no stale buffer, mask or statistic may leak between shapes."""
    import copy
    from gpu_util import build_ours, masks_to_ours
    from oracle import synth
    from oracle import trimodal_oracle as O
    from train_eval import train_gan as TG
    cfg = O.HotPathConfig(n_words=300, n_speakers=12)

    def step(G, D, g_opt, d_opt, B, seed):
        inp, noise = synth.make_inputs(cfg, B, seed=seed), synth.make_noise(cfg, B, seed=seed + 1, dropout=True)
        TG.inject_noise(TG.StepNoise(eps=list(noise.eps), perm=noise.perm, g_masks=[masks_to_ours(m, CPU) for m in noise.g_masks],
                                     d_masks=[masks_to_ours(m, CPU) for m in noise.d_masks]))
        return TG.train_iter_gan(args, 11, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)

    def opts(G, D):
        return (torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999)),
                torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999)))

    args, G, D, _, _ = build_ours(cfg, CPU)
    G.train(); D.train()
    g_opt, d_opt = opts(G, D)
    step(G, D, g_opt, d_opt, 5, 10)
    G.eval()
    inp = synth.make_inputs(cfg, 2, seed=30)
    with torch.no_grad():
        G(O.make_pre_seq(inp['target'], cfg.n_pre_poses), inp['in_text'], inp['in_audio'], inp['vid'])
    G.train()
    g_sd, d_sd = copy.deepcopy(G.state_dict()), copy.deepcopy(D.state_dict())
    a = step(G, D, g_opt, d_opt, 3, 20)
    ga = {k: p.grad.clone() for k, p in G.named_parameters()}
    args, G2, D2, _, _ = build_ours(cfg, CPU)
    G2.load_state_dict(g_sd); D2.load_state_dict(d_sd)
    G2.train(); D2.train()
    b = step(G2, D2, *opts(G2, D2), 3, 20)
    # D's Adam state differs between the two runs (second vs first update), which changes D(G(x)) in the G step a little: compare what
    # does not depend on it tightly (the regression loss) and the rest loosely
    assert abs(a['loss'] - b['loss']) <= 1e-6 * abs(b['loss']) and abs(a['KLD'] - b['KLD']) <= 1e-6 * abs(b['KLD'])
    assert abs(a['dis'] - b['dis']) <= 1e-5 * abs(b['dis'])
    assert abs(a['gen'] - b['gen']) <= 5e-2 * abs(b['gen'])


def test_reference_style_autograd_loop_does_not_accumulate_stale_gradients(emu_fp32):
    """The reference's own training pattern on the nn.Module API: optim.zero_grad() (which DROPS every .grad by default), forward,
    loss.backward(), optim.step().  The .grad views into the flat arena are re-created on the next forward and must start from zero: with
    a zero learning rate two iterations give identical gradients (they were doubled before the fix in ParamArena.bind_grads)."""
    from gpu_util import build_ours
    from conftest import rel_l2
    from oracle import synth
    from oracle import trimodal_oracle as O
    cfg = O.HotPathConfig(n_words=300, n_speakers=12, dropout_prob=0.0, emb_dropout=0.0)
    args, G, D, _, _ = build_ours(cfg, CPU, dropout_prob=0.0)
    G.text_encoder.emb_dropout = 0.0
    G.train(); D.train()
    opt = torch.optim.SGD(list(G.parameters()) + list(D.parameters()), lr=0.0)
    inp = synth.make_inputs(cfg, 3, seed=1)
    pre = O.make_pre_seq(inp['target'], cfg.n_pre_poses)
    eps = synth.make_noise(cfg, 3, seed=1).eps[0]
    grads = []
    for it in range(3):
        opt.zero_grad()
        G.set_noise(eps=eps, masks={})
        D.set_noise(masks={})
        poses, z, mu, logvar = G(pre, inp['in_text'], inp['in_audio'], inp['vid'])
        loss = (poses - inp['target']).abs().mean() + 0.1 * mu.pow(2).mean() - torch.log(D(poses) + 1e-8).mean()
        loss.backward()
        opt.step()
        grads.append({k: p.grad.clone() for m in (G, D) for k, p in m.named_parameters()})
    for k, g0 in grads[0].items():
        if g0.norm() > 1e-8:
            assert rel_l2(grads[1][k], g0) < 1e-5 and rel_l2(grads[2][k], g0) < 1e-5, k


def run_train_variant(dev, ctx, zm, tol=1e-4):
    """One train_iter_gan step (epoch 11, dropout off) of OUR modules built with the other constructor variants / args.z_type values vs the
    reference's own run of the same step (tests/golden/train_variants.npz, oracle/make_golden_variants.py)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from gpu_util import make_args
    from model import vocab
    from model.multimodal_context_net import ConvDiscriminator, PoseGenerator
    from oracle import synth
    from oracle.make_golden import digest, golden_cfg
    from train_eval import train_gan as TG
    cfg = golden_cfg()
    g = np.load(os.path.join(GOLDEN, 'train_variants.npz'))
    tag = f'{ctx}_{zm}'
    args = make_args(cfg)
    args.input_context, args.z_type = ctx, {'speaker': 'speaker', 'random': 'random', None: 'none'}[zm]
    z_obj = None
    if zm == 'speaker':
        z_obj = vocab.Vocab('vid', insert_default_tokens=False)
        while z_obj.n_words < cfg.n_speakers:
            z_obj.index_word('spk%d' % z_obj.n_words)
    elif zm == 'random':
        z_obj = 1
    G = PoseGenerator(args, cfg.pose_dim, cfg.n_words, cfg.wordembed_dim, None, z_obj=z_obj)
    D = ConvDiscriminator(cfg.pose_dim)
    G.load_state_dict(synth.with_tcn_aliases(synth.generator_state_dict_variant(cfg, ctx, zm)), strict=True)
    D.load_state_dict(synth.discriminator_state_dict(cfg), strict=True)
    G, D = G.to(dev).train(), D.to(dev).train()
    inp = {k: v.to(dev) for k, v in synth.make_inputs(cfg, 3, seed=1).items()}
    noise = synth.golden_noise(cfg, 3, 2, False)
    g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
    TG.inject_noise(TG.StepNoise(eps=[e.to(dev) for e in noise.eps], perm=noise.perm.to(dev), g_masks=[{}, {}, {}], d_masks=[{}, {}, {}]))
    ret = TG.train_iter_gan(args, 11, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'] if zm == 'speaker' else None, G, D, g_opt, d_opt)
    want = {k[len(tag) + 6:]: float(g[k]) for k in g.files if k.startswith(tag + '/loss_')}
    assert set(ret) == set(want), (ret, want)
    for k, v in ret.items():
        assert abs(v - want[k]) <= tol * abs(want[k]) + 1e-6, (tag, k, v, want[k])
    for k, p in G.named_parameters():
        if not bool(g[f'{tag}/hasgrad/{k}']) or k in GP.ZERO_GRAD_KEYS:
            continue
        ref = g[f'{tag}/ggrad/{k}']
        GP._digest_close(digest(p.grad.cpu())[:34], ref, 3 * tol)


@pytest.mark.parametrize('ctx,zm', [('audio', 'speaker'), ('text', 'random'), ('none', None), ('both', 'random'), ('none', 'speaker')])
def test_train_iter_gan_constructor_variants(emu_fp32, ctx, zm):
    run_train_variant(CPU, ctx, zm)


@pytest.mark.parametrize('at', ['top', 'concat', 'gru0', 'gru2', 'pre1'])
def test_discriminator_real_pass_placements_are_equivalent(emu_fast, monkeypatch, at):
    """train_iter_gan forks D(real) from inside the generator's forward by default ('gru1'); every other placement (config.d_real_at) is
    the same arithmetic - BatchNorm statistics still see real before fake, D's gradients are the same sum - and meets the same goldens."""
    from tgb200 import config
    monkeypatch.setattr(config, '_DREAL_AT', at)
    GP.test_fast_mode_train_iter_full_size_vs_oracle(CPU, 4, 11, 2000, 50)
