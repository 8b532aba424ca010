"""EmbeddingSpaceEvaluator - Frechet Gesture Distance (reference: scripts/model/embedding_space_evaluator.py:16-156).

push_samples runs the EmbeddingNet kernels on the real and the generated clips and folds the 32-d features straight
into fp64 sufficient statistics on the device (count, sum x, sum x x^T, sum |real-gen|), so nothing is copied to the host
per batch (the reference does 2 x .cpu().numpy() + 2 x .item() per batch, :55-60).  get_scores makes ONE device->host
copy, optionally (reduce=True) sums the statistics of the ranks' shards first, forms mean / covariance (ddof=1, == np.cov) in
float64 and evaluates the 32x32 matrix square root on the host with SciPy exactly like the reference (:138-156)."""
import numpy as np
import torch
from scipy import linalg

from model.embedding_net import EmbeddingNet
from tgb200 import ops


class EmbeddingSpaceEvaluator:
    F = 32

    def __init__(self, args, embed_net_path, lang_model, device):
        self.n_pre_poses = args.n_pre_poses
        ckpt = torch.load(embed_net_path, map_location=device, weights_only=False)
        n_frames = args.n_poses
        word_embeddings = lang_model.word_embedding_weights
        self.pose_dim = ckpt['pose_dim']
        self.net = EmbeddingNet(args, self.pose_dim, n_frames, lang_model.n_words, args.wordembed_dim, word_embeddings, 'pose').to(device)
        self.net.load_state_dict(ckpt['gen_dict'])
        self.net.train(False)
        self.device = device
        self.reset()

    @classmethod
    def from_net(cls, net, n_pre_poses, device):
        """Builds an evaluator around an already-constructed EmbeddingNet (no checkpoint file)."""
        self = cls.__new__(cls)
        self.n_pre_poses, self.net, self.device = n_pre_poses, net.to(device), device
        self.net.train(False)
        self.reset()
        return self

    def reset(self):
        n = 1 + self.F + self.F * self.F
        self.acc_real = torch.zeros(n, dtype=torch.float64, device=self.device)
        self.acc_gen = torch.zeros(n, dtype=torch.float64, device=self.device)
        self.acc_misc = torch.zeros(4, dtype=torch.float64, device=self.device)   # sum|real-gen| feats, recon L1 real, recon L1 gen
        self.n_push = 0
        self.n_elems = 0
        self.context_feat_list = []

    def get_no_of_samples(self):
        return self.n_push

    def push_samples(self, context_text, context_spec, generated_poses, real_poses):
        eng = self.net.engine().ensure(real_poses.device)
        B = real_poses.shape[0]
        real = real_poses.detach().contiguous().float()
        gen = generated_poses.detach().contiguous().float()
        rf, _, _, rrec = eng.forward(real, decode=True)
        ops.feature_stats(rf, B, self.F, self.acc_real)
        ops.l1_dist(real, rrec, real.numel(), self.acc_misc[1:])
        rf_keep = eng.ws.get('eval.rf', (B, self.F))
        rf_keep.copy_(rf)
        gf, _, _, grec = eng.forward(gen, decode=True)
        ops.feature_stats(gf, B, self.F, self.acc_gen)
        ops.l1_dist(gen, grec, gen.numel(), self.acc_misc[2:])
        ops.l1_dist(rf_keep, gf, B * self.F, self.acc_misc)
        self.n_push += 1
        self.n_elems += real.numel()

    @staticmethod
    def _moments(acc, F):
        n = acc[0]
        mu = acc[1:1 + F] / n
        sxx = acc[1 + F:].reshape(F, F)
        cov = (sxx - n * np.outer(mu, mu)) / (n - 1.0)
        return mu, cov

    def get_scores(self, reduce=False):
        """reduce=False (default, like the reference's purely local getter, :74-101): scores of what THIS process pushed.
        reduce=True: every rank of the default process group must call it, each having pushed its own SHARD of the test set; the fp64
        sufficient statistics are summed across ranks first (one all-reduce of 2 x 1057 + 4 doubles), so all ranks return the scores of
        the union.  Do not use it when every rank pushed the full set (the counts would be multiplied by the world size)."""
        acc = torch.stack([self.acc_real, self.acc_gen])
        misc = self.acc_misc.clone()
        if reduce:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(acc)
                dist.all_reduce(misc)
        acc = acc.cpu().numpy()
        misc = misc.cpu().numpy()
        b_mu, b_sigma = self._moments(acc[0], self.F)      # real
        a_mu, a_sigma = self._moments(acc[1], self.F)      # generated
        try:
            frechet_dist = self.calculate_frechet_distance(a_mu, a_sigma, b_mu, b_sigma)
        except ValueError:
            frechet_dist = 1e+10
        feat_dist = float(misc[0] / acc[0][0])
        return frechet_dist, feat_dist

    def recon_err_diff_mean(self):
        """mean over pushed batches of (recon L1 of generated - recon L1 of real), cf. :58-61 (equal batch sizes)."""
        misc = self.acc_misc.cpu().numpy()
        return float((misc[2] - misc[1]) / max(self.n_elems, 1))

    @staticmethod
    def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
        """d^2 = |mu1-mu2|^2 + Tr(S1 + S2 - 2 sqrt(S1 S2)) (embedding_space_evaluator.py:103-156), float64 on the host."""
        mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
        sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
        assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
        diff = mu1 - mu2
        covmean = linalg.sqrtm(sigma1.dot(sigma2))
        if isinstance(covmean, tuple):
            covmean = covmean[0]
        if not np.isfinite(covmean).all():
            offset = np.eye(sigma1.shape[0]) * eps
            covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
            if isinstance(covmean, tuple):
                covmean = covmean[0]
        if np.iscomplexobj(covmean):
            if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
                raise ValueError('Imaginary component {}'.format(np.max(np.abs(covmean.imag))))
            covmean = covmean.real
        return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))
