"""train_iter / evaluate_testset - drop-in for scripts/train_feature_extractor.py:25-97: training of the pose auto-encoder whose
latent is the FGD feature space.  Same signatures and returned dicts; the dataset / checkpoint / video plumbing of the
reference's main() (:100-233, Human3.6M loader) is outside the hot path.  The arithmetic lives in
train_eval.train_joint_embed.ae_step (one CUDA-graph-replayed launch sequence per step)."""
import logging
import time

import torch

from train_eval.train_joint_embed import ae_step, eval_embed

VARIATIONAL_ENCODING = False        # train_feature_extractor.py:58 ("AE or VAE": the reference hard-codes the AE branch)
RECON_WEIGHT = 1                    # :86


def train_iter(args, epoch, target_data, net, optim):
    """train_feature_extractor.py:54-97, AE branch: recon_loss = sum_b [mean|recon - target| + mean|frame differences|] (:64-72),
    loss = 1 * recon_loss (:85-87), backward, Adam step; returns {'loss': recon_weight * recon_loss}."""
    return {'loss': RECON_WEIGHT * ae_step(net, optim, target_data, use_diff=True, weight=float(RECON_WEIGHT))}


def evaluate_testset(test_data_loader, generator):
    """train_feature_extractor.py:25-51: sample-weighted mean of eval_embed's loss in eval mode; the per-batch `.item()` of the
    reference becomes one device accumulator and a single read-back."""
    generator.train(False)
    start = time.time()
    total = count = None
    with torch.no_grad():
        for data in test_data_loader:
            _, target_vec = data
            target = target_vec if target_vec.is_cuda else target_vec.to(next(generator.parameters()).device, non_blocking=True)
            loss, _ = eval_embed(None, None, None, target, generator)
            n = target.shape[0]
            total = loss.double() * n if total is None else total + loss.double() * n
            count = n if count is None else count + n
    generator.train(True)
    avg = float(total.cpu()) / count if count else 0.0
    logging.info('[VAL] loss: {:.3f} / {:.1f}s'.format(avg, time.time() - start))
    return {'loss': avg}
