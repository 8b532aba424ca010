"""evaluate_testset - the validation loop around the hot path (scripts/train.py:234-329) with everything kept on the device.

Same signature and returned dict ({'loss', 'joint_mae'[, 'frechet', 'feat_dist']}) as the reference.  Per batch the reference does a
`.item()`, two `.cpu().numpy()` copies of the poses, a NumPy kinematic-tree walk and two AverageMeter updates; here the generated
poses go straight from the generator into the EmbeddingNet statistics (EmbeddingSpaceEvaluator.push_samples) and into one metrics
kernel (tg_pose_eval_metrics: L1, joint MAE, accel in fp64 accumulators), and the host reads 3 doubles once at the end."""
import logging
import random
import time

import torch

from tgb200 import _lib, ops


class PoseMetrics:
    """Device-side accumulators of the three per-batch meters of evaluate_testset (train.py:240-242,283,293-310).  AverageMeter weights
    each batch mean by its batch size, which for equal-length clips is the global mean computed here."""

    def __init__(self, device):
        self.acc = torch.zeros(3, dtype=torch.float64, device=device)
        self.n_l1 = self.n_mae = self.n_acc = 0

    def push(self, out_dir_vec, target, n_pre_poses):
        B, T, D = target.shape
        out = out_dir_vec.detach().contiguous().float()
        tgt = target.detach().contiguous().float()
        assert out.shape == tgt.shape, 'generators that emit only the non-seed frames (train.py:302-303) are not on the configured path'
        ops.pose_eval_metrics(out, tgt, B, T, D, n_pre_poses, self.acc)
        self.n_l1 += B * T * D
        self.n_mae += B * (T - n_pre_poses) * 30
        self.n_acc += B * (T - 2) * 30

    def result(self):
        a = self.acc.cpu().tolist()
        return {'loss': a[0] / max(self.n_l1, 1), 'joint_mae': a[1] / max(self.n_mae, 1), 'accel': a[2] / max(self.n_acc, 1)}


def evaluate_testset(test_data_loader, generator, loss_fn, embed_space_evaluator, args):
    """train.py:234-329 for args.model in {'multimodal_context', 'seq2seq', 'speech2gesture', 'joint_embedding', 'gesture_autoencoder'}.  `loss_fn` is
    accepted for signature compatibility; like the reference's multimodal_context branch the reported loss is the L1 distance of the
    direction vectors (for the two embedding models that IS eval_embed's loss: the batch mean of per-sample means over equal-sized samples)."""
    _lib.require_cuda()
    was_training = generator.training
    generator.train(False)
    if embed_space_evaluator:
        embed_space_evaluator.reset()
    start = time.time()
    metrics = None
    gen = generator.module if hasattr(generator, 'module') else generator
    with torch.no_grad():
        for data in test_data_loader:
            in_text, text_lengths, in_text_padded, _, target_vec, in_audio, in_spec, aux_info = data
            dev = next(gen.parameters()).device
            target = target_vec.to(dev, non_blocking=True)
            batch_size = target.size(0)
            if metrics is None:
                metrics = PoseMetrics(dev)
            model = getattr(args, 'model', 'multimodal_context')
            if model == 'multimodal_context':
                speaker_model = getattr(gen, 'z_obj', None)
                vid_indices = None
                if speaker_model is not None and hasattr(speaker_model, 'word2index'):
                    ids = list(speaker_model.word2index.values())
                    vid_indices = torch.LongTensor([random.choice(ids) for _ in range(batch_size)]).to(dev)
                pre_seq = target.new_zeros((target.shape[0], target.shape[1], target.shape[2] + 1))
                pre_seq[:, 0:args.n_pre_poses, :-1] = target[:, 0:args.n_pre_poses]
                pre_seq[:, 0:args.n_pre_poses, -1] = 1
                out_dir_vec, *_ = generator(pre_seq, in_text_padded.to(dev), in_audio.to(dev), vid_indices)
            elif model == 'seq2seq':
                out_dir_vec = generator(in_text.to(dev), text_lengths, target, None)
            elif model == 'speech2gesture':                                     # train.py:277-279
                out_dir_vec = generator(in_spec.to(dev), target[:, 0:args.n_pre_poses])
            elif model == 'joint_embedding':                                    # train.py:269-271: decode the speech latent
                from train_eval.train_joint_embed import eval_embed
                _, out_dir_vec = eval_embed(in_text_padded.to(dev), in_audio.to(dev), target[:, 0:args.n_pre_poses], target, generator, mode='speech')
            elif model == 'gesture_autoencoder':                                # train.py:272-273,279: loss only, no pose metrics / FGD
                from train_eval.train_joint_embed import eval_embed
                _, recon = eval_embed(None, None, target[:, 0:args.n_pre_poses], target, generator)
                metrics.push(recon, target, args.n_pre_poses)
                continue
            else:
                raise _lib.TgError('evaluate_testset: model %r is not on the B200 path' % model)
            if embed_space_evaluator:
                embed_space_evaluator.push_samples(in_text_padded, in_audio, out_dir_vec, target)
            metrics.push(out_dir_vec, target, args.n_pre_poses)
    generator.train(was_training)
    res = metrics.result() if metrics is not None else {'loss': 0.0, 'joint_mae': 0.0, 'accel': 0.0}
    if getattr(args, 'model', None) == 'gesture_autoencoder':
        res['joint_mae'] = res['accel'] = 0.0                                   # never updated in the reference (train.py:279): AverageMeter.avg == 0
    ret_dict = {'loss': res['loss'], 'joint_mae': res['joint_mae']}
    elapsed_time = time.time() - start
    if embed_space_evaluator and embed_space_evaluator.get_no_of_samples() > 0:
        frechet_dist, feat_dist = embed_space_evaluator.get_scores()
        logging.info('[VAL] loss: {:.3f}, joint mae: {:.5f}, accel diff: {:.5f}, FGD: {:.3f}, feat_D: {:.3f} / {:.1f}s'.format(
            res['loss'], res['joint_mae'], res['accel'], frechet_dist, feat_dist, elapsed_time))
        ret_dict['frechet'] = frechet_dist
        ret_dict['feat_dist'] = feat_dist
    else:
        logging.info('[VAL] loss: {:.3f}, joint mae: {:.3f} / {:.1f}s'.format(res['loss'], res['joint_mae'], elapsed_time))
    return ret_dict
