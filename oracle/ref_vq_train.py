"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's training-mode VQ codebook
(modules/quantization/core_vq.py).  Only tests/ may import this; the product never does.

Pinned: oracle/make_golden_vq_train.py runs the REAL reference (EuclideanCodebook / VectorQuantization in train mode,
three steps, k-means initialisation, a dead-code expiry) with its random draws recorded, asserts that this restatement
reproduces every buffer bit for bit from the same draws, and writes tests/golden/vq_train.npz.

Random draws are explicit arguments here (the reference calls torch.randperm / torch.randint inside sample_vectors,
core_vq.py:63-71)."""
import torch


def kmeans(samples, num_clusters, num_iters, init_indices):
    """core_vq.py:74-96 -> (means, bins of the last iteration)."""
    dim, dtype = samples.shape[-1], samples.dtype
    means = samples[init_indices]                                            # :77 sample_vectors
    bins = None
    for _ in range(num_iters):
        diffs = samples[:, None, :] - means[None, :, :]                      # :80-82
        dists = -(diffs ** 2).sum(dim=-1)                                    # :83
        buckets = dists.max(dim=-1).indices                                  # :85
        bins = torch.bincount(buckets, minlength=num_clusters)               # :86
        zero_mask = bins == 0
        bins_min_clamped = bins.masked_fill(zero_mask, 1)                    # :88
        new_means = buckets.new_zeros(num_clusters, dim, dtype=dtype)
        new_means.scatter_add_(0, buckets[:, None].expand(-1, dim), samples)  # :91
        new_means = new_means / bins_min_clamped[..., None]                  # :92
        means = torch.where(zero_mask[..., None], means, new_means)          # :94
    return means, bins


def quantize(x, embed):
    """core_vq.py:175-183"""
    e = embed.t()
    dist = -(x.pow(2).sum(1, keepdim=True) - 2 * x @ e + e.pow(2).sum(0, keepdim=True))
    return dist.max(dim=-1).indices


class Codebook:
    """Buffers + train-mode forward of EuclideanCodebook (core_vq.py:100-231)."""

    def __init__(self, dim, codebook_size, kmeans_iters=10, decay=0.99, epsilon=1e-5, threshold_ema_dead_code=2,
                 embed=None):
        self.K, self.kmeans_iters, self.decay, self.eps, self.thr = codebook_size, kmeans_iters, decay, epsilon, threshold_ema_dead_code
        self.inited = embed is not None
        self.embed = embed.clone() if embed is not None else torch.zeros(codebook_size, dim)
        self.embed_avg = self.embed.clone()
        self.cluster_size = torch.zeros(codebook_size)

    def forward_train(self, x, init_indices=None, expire_pick=None):
        """x (..., D); returns (quantize, ind).  init_indices: the draw of the k-means start (first call only);
        expire_pick: the draw used when some cluster_size < threshold (core_vq.py:158-169)."""
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        if not self.inited:                                                  # :141-149
            embed, cs = kmeans(x, self.K, self.kmeans_iters, init_indices)
            self.embed = embed.clone()
            self.embed_avg = embed.clone()
            self.cluster_size = cs.to(torch.float32)
            self.inited = True
        ind = quantize(x, self.embed)                                        # :209
        onehot = torch.nn.functional.one_hot(ind, self.K).type(x.dtype)
        q = torch.nn.functional.embedding(ind.view(*shape[:-1]), self.embed)  # :212
        expired = self.cluster_size < self.thr                               # :162
        used_pick = False
        if self.thr != 0 and bool(expired.any()):
            self.embed = torch.where(expired[..., None], x[expire_pick], self.embed)   # :151-156
            used_pick = True
        self.cluster_size = self.cluster_size.clone().mul_(self.decay).add_(onehot.sum(0), alpha=(1 - self.decay))   # :219 (ema_inplace :48-49)
        embed_sum = x.t() @ onehot                                           # :220
        self.embed_avg = self.embed_avg.clone().mul_(self.decay).add_(embed_sum.t(), alpha=(1 - self.decay))     # :221
        n = self.cluster_size.sum()
        cs = (self.cluster_size + self.eps) / (n + self.K * self.eps) * n    # :222-226 laplace_smoothing * sum
        self.embed = self.embed_avg / cs.unsqueeze(1)                        # :227-228
        return q, ind.view(*shape[:-1]), used_pick


def vq_forward_train(cb: Codebook, x_bdn, commitment_weight=1.0, **draws):
    """VectorQuantization.forward in train mode (core_vq.py:294-316): x (B, D, N) requires_grad ->
    (quantize (B, D, N) with the straight-through graph, ind (B, N), loss (1,), used_pick)."""
    x = x_bdn.transpose(1, 2)
    q, ind, used = cb.forward_train(x.detach(), **draws)
    q = x + (q - x).detach()                                                 # :301
    loss = torch.tensor([0.0], requires_grad=True)
    if commitment_weight > 0:
        loss = loss + torch.nn.functional.mse_loss(q.detach(), x) * commitment_weight   # :309-310
    return q.transpose(1, 2), ind, loss, used
