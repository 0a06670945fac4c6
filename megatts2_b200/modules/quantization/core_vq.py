"""EuclideanCodebook / VectorQuantization / ResidualVectorQuantization with the reference's
surface (modules/quantization/core_vq.py:100-231, 234-316, 319-367): same ctor kwargs, module
tree and buffers (``_codebook.{inited,cluster_size,embed,embed_avg}``), same
forward/encode/decode signatures.  The eval-time search and gather run in
libmegatts2_b200 (mtts_vq_argmin_f32 / mtts_vq_gather_f32).  Train mode adds the
reference's k-means initialisation, dead-code expiry, EMA codebook update and buffer
broadcast (core_vq.py:141-169, 217-229) and the straight-through / commitment step
(:300-311) through megatts2_b200/vq_train.py."""
import typing as tp

import torch
from torch import nn

from ... import _lib as L
from ... import ops
from ... import vq_train as VT


class EuclideanCodebook(nn.Module):
    def __init__(self, dim: int, codebook_size: int, kmeans_init: int = False, kmeans_iters: int = 10,
                 decay: float = 0.99, epsilon: float = 1e-5, threshold_ema_dead_code: int = 2):
        super().__init__()
        self.decay = decay
        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.epsilon = epsilon
        self.threshold_ema_dead_code = threshold_ema_dead_code
        embed = torch.zeros(codebook_size, dim) if kmeans_init else _uniform_init(codebook_size, dim)
        self.register_buffer("inited", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())
        self._inited_sig = None

    def _is_inited(self) -> bool:
        sig = (self.inited.data_ptr(), self.inited._version)
        if self._inited_sig != sig:     # one readback per (re)load, not per call
            self._inited_val = bool(self.inited.detach().cpu().item())
            self._inited_sig = sig
        return self._inited_val

    def _require_inference(self):
        if not self._is_inited():
            raise L.MttsError("codebook not initialised (inited == 0): the reference would run k-means here "
                              "(core_vq.py:141-149); call forward() in train mode or load a trained codebook first")

    def init_embed_(self, data, init_indices=None):
        """k-means initialisation on the first training batch (core_vq.py:141-149)."""
        if self._is_inited():
            return
        embed, cluster_size = VT.kmeans(data, self.codebook_size, self.kmeans_iters, init_indices)
        self.embed.data.copy_(embed)
        self.embed_avg.data.copy_(embed)
        self.cluster_size.data.copy_(cluster_size)
        self.inited.data.fill_(1.0)
        self._inited_sig = None            # .data writes do not bump the version the cache is keyed on
        VT.broadcast_buffers(self.buffers())

    def expire_codes_(self, batch_samples, pick=None):
        """Dead-code expiry (core_vq.py:158-169).  The reference returns early (and skips the broadcast) when no code is
        below the threshold, after a host-side ``torch.any``; here the masked replacement always runs on the device
        (a no-op without dead codes) and the buffers are broadcast every step, which costs no host sync."""
        if self.threshold_ema_dead_code == 0:
            return
        VT.replace_expired(self.embed, batch_samples.reshape(-1, batch_samples.shape[-1]), self.cluster_size,
                           self.threshold_ema_dead_code, pick)
        VT.broadcast_buffers(self.buffers())

    def quantize(self, x):
        """x (N, D) -> (N,) int64: first index of max_k -(|x|^2 - 2 x.e_k + |e_k|^2) (core_vq.py:175-183)."""
        return ops.vq_argmin(x, self.embed)

    def dequantize(self, embed_ind):
        shp = embed_ind.shape
        return ops.vq_gather(embed_ind.reshape(1, -1), self.embed).reshape(*shp, self.embed.shape[1])

    def encode(self, x):
        self._require_inference()
        shape = x.shape
        flat = x.reshape(-1, shape[-1])
        return self.quantize(flat).view(*shape[:-1])

    def decode(self, embed_ind):
        return self.dequantize(embed_ind)

    def forward(self, x):
        shape = x.shape
        flat = x.detach().reshape(-1, shape[-1])
        if self.training:
            self.init_embed_(flat)
        else:
            self._require_inference()
        ind_flat = self.quantize(flat)
        ind = ind_flat.view(*shape[:-1])
        quantize = self.dequantize(ind)
        if self.training:
            # expiry first, then the EMA update from this batch's assignment (core_vq.py:214-229)
            self.expire_codes_(flat)
            sums, counts = VT.cluster_sum(flat, ind_flat, self.codebook_size)
            VT.ema_update(self.cluster_size, self.embed_avg, self.embed, sums, counts, self.decay, self.epsilon)
        return quantize, ind


def _uniform_init(*shape):
    t = torch.empty(shape)
    nn.init.kaiming_uniform_(t)
    return t


class VectorQuantization(nn.Module):
    def __init__(self, dim: int, codebook_size: int, codebook_dim: tp.Optional[int] = None, decay: float = 0.99,
                 epsilon: float = 1e-5, kmeans_init: bool = True, kmeans_iters: int = 50,
                 threshold_ema_dead_code: int = 2, commitment_weight: float = 1.):
        super().__init__()
        _codebook_dim = codebook_dim if codebook_dim is not None else dim
        if _codebook_dim != dim:
            raise L.MttsError("projected codebooks (codebook_dim != dim) are not on the synthesis path")
        self.project_in = nn.Identity()
        self.project_out = nn.Identity()
        self.epsilon = epsilon
        self.commitment_weight = commitment_weight
        self._codebook = EuclideanCodebook(dim=_codebook_dim, codebook_size=codebook_size, kmeans_init=kmeans_init,
                                           kmeans_iters=kmeans_iters, decay=decay, epsilon=epsilon,
                                           threshold_ema_dead_code=threshold_ema_dead_code)
        self.codebook_size = codebook_size

    @property
    def codebook(self):
        return self._codebook.embed

    def encode(self, x):
        """x (B, D, N) -> (B, N) int64"""
        return self._codebook.encode(ops.to_channels_last(x))

    def decode(self, embed_ind):
        """(B, N) int64 -> (B, D, N)"""
        return ops.to_channels_first(self._codebook.decode(embed_ind))

    def forward(self, x):
        """VectorQuantization.forward (core_vq.py:294-316): (quantize (B,D,N), ind (B,N), loss [1]).  Train mode:
        straight-through output x + (q - x).detach() and loss = commitment_weight * mse(quantize.detach(), x)."""
        if not self.training:
            q, ind = self._codebook(ops.to_channels_last(x))
            return ops.to_channels_first(q), ind, torch.zeros(1, device=x.device)
        xl = VT.channels_last(x)
        q, ind = self._codebook(xl)
        q, mse = VT.StraightThroughCommit.apply(xl, q, self.commitment_weight)
        loss = torch.zeros(1, device=x.device, requires_grad=True)
        if self.commitment_weight > 0:
            loss = loss + mse * self.commitment_weight
        return VT.channels_first(q), ind, loss


class ResidualVectorQuantization(nn.Module):
    def __init__(self, *, num_quantizers, **kwargs):
        super().__init__()
        if num_quantizers != 1:
            raise L.MttsError("the synthesis path uses n_q == 1 (modules/vqpe.py:44-49)")
        self.layers = nn.ModuleList([VectorQuantization(**kwargs) for _ in range(num_quantizers)])

    def forward(self, x, n_q: tp.Optional[int] = None):
        quantized, indices, loss = self.layers[0](x)
        return quantized, indices.unsqueeze(0), loss.unsqueeze(0)

    def encode(self, x: torch.Tensor, n_q: tp.Optional[int] = None) -> torch.Tensor:
        return self.layers[0].encode(x).unsqueeze(0)

    def decode(self, q_indices: torch.Tensor) -> torch.Tensor:
        assert q_indices.shape[0] == 1
        return self.layers[0].decode(q_indices[0])
