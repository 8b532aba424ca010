"""CPU oracle for the seq2seq baseline (config/seq2seq.yml).  TEST INFRASTRUCTURE ONLY (same rules as trimodal_oracle.py).

Functional torch restatement (fp32 or fp64) of
  * Seq2SeqNet.forward                scripts/model/seq2seq_net.py:217-254
      - EncoderRNN.forward            :38-59   (embedding -> packed 2-layer bi-GRU -> sum of directions)
      - Attn.forward / score          :72-94   (tanh(W [h; enc]) . v -> softmax over ALL padded positions -> context)
      - BahdanauAttnDecoderRNN.forward:150-198 (cat(input, context) -> Linear -> BatchNorm1d -> ReLU -> 2-layer GRU step -> Linear)
  * custom_loss / train_iter_seq2seq  scripts/train_eval/train_seq2seq.py:6-51 (MSE + continuity + variance terms,
    clip_grad_norm_(5), Adam).
Weights are a plain dict with the reference's state_dict keys.  Packed-sequence semantics are restated with length masks:
a sample's forward chain stops updating at its length (its final hidden state is the state at its last valid step, padded
outputs are zero); its reverse chain starts, from h = 0, at its last valid step.

Parity pinning: the reference has no tests for this path; oracle/make_golden_seq2seq.py runs the reference modules
(imported from /root/reference in the build container) and writes tests/golden/seq2seq_step.npz, against which
tests/test_oracle_golden.py re-checks this file on every CPU run."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class Seq2SeqConfig:
    """config/seq2seq.yml:7,17-19,25-28,38-39 + train.py:497 (pose_dim)."""
    n_words: int = 20000
    wordembed_dim: int = 300
    hidden_size: int = 200
    n_layers: int = 2
    dropout_prob: float = 0.1
    pose_dim: int = 27
    n_poses: int = 34
    n_pre_poses: int = 4
    learning_rate: float = 1e-4
    loss_regression_weight: float = 250.0
    loss_kld_weight: float = 0.1          # continuity-term weight (train_seq2seq.py:19)
    loss_reg_weight: float = 25.0         # variance-term weight   (train_seq2seq.py:24)
    max_grad_norm: float = 5.0            # train_seq2seq.py:48
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1


def gru_cell(x_gi: Tensor, h: Tensor, w_hh: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU step from the pre-computed input projection gi = W_ih x + b_ih (gate order r,z,n)."""
    H = h.shape[1]
    gh = h @ w_hh.t() + b_hh
    i_r, i_z, i_n = x_gi.split(H, dim=1)
    h_r, h_z, h_n = gh.split(H, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def packed_gru_direction(x: Tensor, lengths: Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> Tuple[Tensor, Tensor]:
    """x [B,T,I] batch-major, lengths [B].  Returns (outputs [B,T,H] with zeros at padded steps, final hidden [B,H]) -
    what nn.GRU returns for a PackedSequence (seq2seq_net.py:54-56)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    outs: List[Optional[Tensor]] = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        valid = (t < lengths).to(x.dtype).unsqueeze(1)
        h_new = gru_cell(gi[:, t], h, w_hh, b_hh)
        h = valid * h_new + (1 - valid) * h
        outs[t] = valid * h_new
    return torch.stack(outs, dim=1), h


def encoder_forward(sd: SD, cfg: Seq2SeqConfig, in_text: Tensor, lengths: Tensor,
                    layer_masks: Optional[List[Optional[Tensor]]] = None) -> Tuple[Tensor, Tensor]:
    """EncoderRNN.forward (seq2seq_net.py:38-59), batch-major.  in_text [B,T] int64 (sorted by decreasing length, padded
    with 0), lengths [B].  Returns (outputs [B,Tmax,H] = fwd + rev, hidden [2*n_layers,B,H] in torch order
    l0 fwd, l0 rev, l1 fwd, l1 rev).  layer_masks[l] multiplies the [B,T,2H] output of layer l < n_layers-1 (train-mode
    inter-layer dropout; padded positions are zero either way)."""
    Tmax = int(lengths.max())
    x = sd['encoder.embedding.weight'][in_text[:, :Tmax]]
    hiddens = []
    inp = x
    for l in range(cfg.n_layers):
        outs = []
        for suffix, rev in (('', False), ('_reverse', True)):
            p = f'encoder.gru.%s_l{l}{suffix}'
            o, h = packed_gru_direction(inp, lengths, sd[p % 'weight_ih'], sd[p % 'weight_hh'], sd[p % 'bias_ih'], sd[p % 'bias_hh'], rev)
            outs.append(o)
            hiddens.append(h)
        inp = torch.cat(outs, dim=2)
        if layer_masks is not None and l < cfg.n_layers - 1 and layer_masks[l] is not None:
            inp = inp * layer_masks[l]
    H = cfg.hidden_size
    return inp[:, :, :H] + inp[:, :, H:], torch.stack(hiddens, dim=0)


def attention(sd: SD, h_last: Tensor, enc: Tensor) -> Tuple[Tensor, Tensor]:
    """Attn.forward/score (seq2seq_net.py:72-94).  h_last [B,H], enc [B,T,H] -> (weights [B,T], context [B,H]).  The softmax
    runs over every padded position too (the reference applies no length mask)."""
    pre = 'decoder.decoder.attn.'
    B, T, H = enc.shape
    cat = torch.cat([h_last.unsqueeze(1).expand(B, T, H), enc], dim=2)
    energy = torch.tanh(cat @ sd[pre + 'attn.weight'].t() + sd[pre + 'attn.bias'])
    score = energy @ sd[pre + 'v']
    w = torch.softmax(score, dim=1)
    return w, torch.bmm(w.unsqueeze(1), enc).squeeze(1)


def decoder_step(sd: SD, cfg: Seq2SeqConfig, x_in: Tensor, hidden: List[Tensor], enc: Tensor, training: bool,
                 bn_state: Optional[Dict[str, Tensor]], layer_mask: Optional[Tensor]) -> Tuple[Tensor, List[Tensor]]:
    """BahdanauAttnDecoderRNN.forward (seq2seq_net.py:150-198) for one time step.  hidden = [h_l0, h_l1] ([B,H] each)."""
    pre = 'decoder.decoder.'
    _, ctx = attention(sd, hidden[-1], enc)
    rnn_in = torch.cat([x_in, ctx], dim=1)
    y = rnn_in @ sd[pre + 'pre_linear.0.weight'].t() + sd[pre + 'pre_linear.0.bias']
    g, b = sd[pre + 'pre_linear.1.weight'], sd[pre + 'pre_linear.1.bias']
    if training:
        mean = y.mean(0)
        var = y.var(0, unbiased=False)
        if bn_state is not None:
            n = y.shape[0]
            bn_state['running_mean'] = (1 - cfg.bn_momentum) * bn_state['running_mean'] + cfg.bn_momentum * mean.detach()
            bn_state['running_var'] = (1 - cfg.bn_momentum) * bn_state['running_var'] + cfg.bn_momentum * var.detach() * n / (n - 1)
            bn_state['num_batches_tracked'] = bn_state['num_batches_tracked'] + 1
    else:
        mean, var = sd[pre + 'pre_linear.1.running_mean'], sd[pre + 'pre_linear.1.running_var']
    y = torch.relu((y - mean) / torch.sqrt(var + cfg.bn_eps) * g + b)
    new_hidden = []
    inp = y
    for l in range(cfg.n_layers):
        p = pre + f'gru.%s_l{l}'
        gi = inp @ sd[p % 'weight_ih'].t() + sd[p % 'bias_ih']
        h = gru_cell(gi, hidden[l], sd[p % 'weight_hh'], sd[p % 'bias_hh'])
        new_hidden.append(h)
        inp = h
        if layer_mask is not None and l < cfg.n_layers - 1:
            inp = inp * layer_mask
    out = inp @ sd[pre + 'out.weight'].t() + sd[pre + 'out.bias']
    return out, new_hidden


def seq2seq_forward(sd: SD, cfg: Seq2SeqConfig, in_text: Tensor, lengths: Tensor, poses: Tensor, training: bool,
                    bn_state: Optional[Dict[str, Tensor]] = None, enc_masks=None, dec_masks=None) -> Tensor:
    """Seq2SeqNet.forward (seq2seq_net.py:229-254).  poses [B,n_poses,D] -> outputs [B,n_poses,D] (frame 0 is copied from
    the input; frames < n_pre_poses are teacher-forced, later frames feed the previous prediction back)."""
    enc, enc_hidden = encoder_forward(sd, cfg, in_text, lengths, enc_masks)
    hidden = [enc_hidden[l] for l in range(cfg.n_layers)]          # encoder_hidden[:n_layers]: l0 fwd, l0 rev (:241)
    outs = [poses[:, 0]]
    x_in = poses[:, 0]
    for t in range(1, cfg.n_poses):
        out, hidden = decoder_step(sd, cfg, x_in, hidden, enc, training, bn_state, dec_masks[t - 1] if dec_masks is not None else None)
        outs.append(out)
        x_in = poses[:, t] if t < cfg.n_pre_poses else out
    return torch.stack(outs, dim=1)


def custom_loss(cfg: Seq2SeqConfig, output: Tensor, target: Tensor) -> Tensor:
    """train_seq2seq.py:6-36.  Note torch.norm(output, 2, 1) reduces over TIME (dim 1), per (sample, joint dim)."""
    n = output.numel()
    mse = ((output - target) ** 2).mean() * cfg.loss_regression_weight
    cont = (output[:, 1:] - output[:, :-1]).abs().sum() / n * cfg.loss_kld_weight
    var = -(output.pow(2).sum(dim=1).sqrt().sum()) / n * cfg.loss_reg_weight
    return mse + cont + var


def is_param(k: str) -> bool:
    return not (k.endswith('running_mean') or k.endswith('running_var') or k.endswith('num_batches_tracked'))


def train_iter_seq2seq_oracle(cfg: Seq2SeqConfig, sd: SD, in_text: Tensor, lengths: Tensor, target: Tensor,
                              opt: Optional[Dict[str, Dict[str, Tensor]]] = None, step: int = 1, betas=(0.9, 0.999), eps: float = 1e-8,
                              enc_masks=None, dec_masks=None):
    """train_iter_seq2seq (train_seq2seq.py:39-51) on a weight dict.  Returns dict(loss, outputs, grads (after clipping),
    total_norm, new_sd (post-Adam weights + BatchNorm running statistics), opt)."""
    dtype = target.dtype
    leaf = {k: (v.detach().clone().to(dtype).requires_grad_(True) if is_param(k) else v.detach().clone()) for k, v in sd.items()}
    bnp = 'decoder.decoder.pre_linear.1.'
    bn_state = {k: leaf[bnp + k] for k in ('running_mean', 'running_var', 'num_batches_tracked')}
    outputs = seq2seq_forward(leaf, cfg, in_text, lengths, target, True, bn_state, enc_masks, dec_masks)
    loss = custom_loss(cfg, outputs, target)
    names = [k for k in leaf if is_param(k)]
    gl = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaf[k])) for k, g in zip(names, gl)}
    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(dtype)
    coef = torch.clamp(cfg.max_grad_norm / (total_norm + 1e-6), max=1.0)            # torch.nn.utils.clip_grad_norm_
    grads = {k: g * coef for k, g in grads.items()}
    if opt is None:
        opt = {k: {'m': torch.zeros_like(leaf[k]), 'v': torch.zeros_like(leaf[k])} for k in names}
    new_sd = {}
    b1, b2 = betas
    for k in leaf:
        if not is_param(k):
            new_sd[k] = bn_state[k[len(bnp):]] if k.startswith(bnp) else leaf[k]
            continue
        g = grads[k]
        m = opt[k]['m'] = b1 * opt[k]['m'] + (1 - b1) * g
        v = opt[k]['v'] = b2 * opt[k]['v'] + (1 - b2) * g * g
        mhat = m / (1 - b1 ** step)
        denom = (v / (1 - b2 ** step)).sqrt() + eps
        new_sd[k] = (leaf[k] - cfg.learning_rate * mhat / denom).detach()
    return dict(loss=loss.detach(), outputs=outputs.detach(), grads=grads, total_norm=total_norm, new_sd=new_sd, opt=opt)
