"""Training-only pieces of the VQ codebook over the C ABI (second half): k-means initialisation
(modules/quantization/core_vq.py:74-96), dead-code expiry (:151-169), the EMA codebook update (:217-229), the
straight-through estimator + commitment loss of VectorQuantization.forward (:294-316) and the buffer broadcast that
keeps the workers' codebooks identical (utils/distrib.py:55-68 -> torch.distributed over NCCL / NVLink).

Random draws (``sample_vectors``: randperm / randint) are torch's, on the samples' device, exactly the calls the
reference makes; every call site accepts explicit indices so that a test can replay the reference's draw."""
import torch

from . import _lib as L
from . import ops

_F32 = torch.float32


def _flat(x, name):
    x = ops._dev(x, name=name)
    return x if x.is_contiguous() else x.contiguous()


def sample_indices(num_samples: int, num: int, device) -> torch.Tensor:
    """Indices sample_vectors draws (core_vq.py:63-71): a random permutation's head when there are enough samples,
    otherwise draws with replacement."""
    if num_samples >= num:
        return torch.randperm(num_samples, device=device)[:num]
    return torch.randint(0, num_samples, (num,), device=device)


def kmeans_assign(samples, means):
    samples, means = _flat(samples, "samples"), _flat(means, "means")
    N, D = samples.shape
    idx = torch.empty(N, dtype=torch.int64, device=samples.device)
    L.check(L.lib().mtts_kmeans_assign_f32(ops._ptr(samples), ops._ptr(means), N, means.shape[0], D, ops._ptr(idx),
                                           ops._stream()))
    return idx


def cluster_sum(samples, idx, K):
    """-> (sum (K, D), count (K,) float32): per-code sums in sample order and bucket sizes."""
    samples = _flat(samples, "samples")
    idx = ops._dev(idx, torch.int64, "idx").contiguous()
    N, D = samples.shape
    s = torch.empty(K, D, dtype=_F32, device=samples.device)
    c = torch.empty(K, dtype=_F32, device=samples.device)
    L.check(L.lib().mtts_vq_cluster_sum_f32(ops._ptr(samples), ops._ptr(idx), N, K, D, ops._ptr(s), ops._ptr(c),
                                            ops._stream()))
    return s, c


def kmeans(samples, num_clusters: int, num_iters: int = 10, init_indices=None):
    """kmeans() of core_vq.py:74-96 -> (means (K, D), bins (K,) int64 of the LAST iteration)."""
    samples = _flat(samples, "samples")
    N, D = samples.shape
    if init_indices is None:
        init_indices = sample_indices(N, num_clusters, samples.device)
    means = samples.index_select(0, init_indices.to(samples.device)).contiguous()
    cnt = torch.zeros(num_clusters, dtype=_F32, device=samples.device)
    for _ in range(num_iters):
        idx = kmeans_assign(samples, means)
        s, cnt = cluster_sum(samples, idx, num_clusters)
        L.check(L.lib().mtts_kmeans_update_f32(ops._ptr(means), ops._ptr(s), ops._ptr(cnt), num_clusters, D, ops._stream()))
    return means, cnt.to(torch.int64)


def ema_update(cluster_size, embed_avg, embed, sums, counts, decay: float, eps: float):
    """In place on the three buffers (core_vq.py:219-229)."""
    for t, n in ((cluster_size, "cluster_size"), (embed_avg, "embed_avg"), (embed, "embed")):
        ops._dev(t, name=n)
        if not t.is_contiguous():
            raise L.MttsError(f"{n}: buffer must be contiguous")
    K, D = embed.shape
    scratch = torch.empty(K, dtype=_F32, device=embed.device)
    L.check(L.lib().mtts_vq_ema_update_f32(ops._ptr(cluster_size), ops._ptr(embed_avg), ops._ptr(embed), ops._ptr(sums),
                                           ops._ptr(counts), K, D, float(decay), float(eps), ops._ptr(scratch), ops._stream()))


def replace_expired(embed, samples, cluster_size, threshold: float, pick=None):
    """expire_codes_ / replace_ (core_vq.py:151-169) without the reference's host-side ``torch.any`` test: rows whose
    EMA bucket size is below the threshold take the sample at pick[k]; the others are untouched."""
    samples = _flat(samples, "samples")
    K, D = embed.shape
    if pick is None:
        pick = sample_indices(samples.shape[0], K, samples.device)
    pick = ops._dev(pick.to(samples.device), torch.int64, "pick").contiguous()
    L.check(L.lib().mtts_vq_replace_rows_f32(ops._ptr(embed), ops._ptr(samples), ops._ptr(pick), ops._ptr(cluster_size),
                                             float(threshold), K, D, samples.shape[0], ops._stream()))


class StraightThroughCommit(torch.autograd.Function):
    """(x, q) -> (x + (q - x), mean((x + (q - x) - x)^2)); gradients reach x only: the pass-through of the quantised
    output plus the commitment term (q is the detached codebook row, core_vq.py:300-311)."""

    @staticmethod
    def forward(ctx, x, q, weight):
        x, q = _flat(x, "x"), _flat(q, "q")
        out = torch.empty_like(x)
        partials = torch.empty(256, dtype=_F32, device=x.device)
        loss = torch.empty(1, dtype=_F32, device=x.device)
        L.check(L.lib().mtts_vq_ste_commit_f32(ops._ptr(x), ops._ptr(q), x.numel(), ops._ptr(out), ops._ptr(partials),
                                               ops._ptr(loss), ops._stream()))
        ctx.save_for_backward(x, out)
        ctx.weight = float(weight)
        return out, loss

    @staticmethod
    def backward(ctx, g_out, g_loss):
        x, out = ctx.saved_tensors
        dx = torch.empty_like(x)
        g_out = None if g_out is None else _flat(g_out, "g_out")
        g_loss = None if g_loss is None else _flat(g_loss, "g_loss")
        # the loss output is the UNWEIGHTED mse; the caller multiplies by commitment_weight, so g_loss carries it
        L.check(L.lib().mtts_vq_ste_commit_bwd_f32(ops._ptr(x), ops._ptr(out), ops._ptr(g_out), ops._ptr(g_loss),
                                                   2.0 / x.numel(), x.numel(), ops._ptr(dx), ops._stream()))
        return dx, None, None


class _ChannelsLast(torch.autograd.Function):
    """(B, C, T) -> contiguous (B, T, C) through the library's tiled copy; the gradient takes the inverse copy."""

    @staticmethod
    def forward(ctx, x):
        return ops.to_channels_last(x)

    @staticmethod
    def backward(ctx, g):
        return ops.to_channels_first(g)


class _ChannelsFirst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.to_channels_first(x)

    @staticmethod
    def backward(ctx, g):
        return ops.to_channels_last(g)


def channels_last(x):
    return _ChannelsLast.apply(x) if x.requires_grad else ops.to_channels_last(x)


def channels_first(x):
    return _ChannelsFirst.apply(x) if x.requires_grad else ops.to_channels_first(x)


def broadcast_buffers(tensors, src: int = 0):
    """distrib.broadcast_tensors (utils/distrib.py:55-68): floating-point buffers of rank ``src`` overwrite everyone
    else's; a no-op outside a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = [t for t in tensors if torch.is_floating_point(t) or torch.is_complex(t)]
    if not tensors:
        return
    count = torch.tensor([len(tensors)], device=tensors[0].device, dtype=torch.long)
    dist.all_reduce(count)
    if int(count.item()) != len(tensors) * dist.get_world_size():
        raise RuntimeError(f"Mismatch in number of params: ours is {len(tensors)}, at least one worker has a different one.")
    handles = [dist.broadcast(t.data, src=src, async_op=True) for t in tensors]
    for h in handles:
        h.wait()
