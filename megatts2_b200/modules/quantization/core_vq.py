"""EuclideanCodebook / VectorQuantization / ResidualVectorQuantization with the reference's
surface (modules/quantization/core_vq.py:100-231, 234-316, 319-367): same ctor kwargs, module
tree and buffers (``_codebook.{inited,cluster_size,embed,embed_avg}``), same
forward/encode/decode signatures.  The eval-time search and gather run in
libmegatts2_b200 (mtts_vq_argmin_f32 / mtts_vq_gather_f32).  Training-only pieces (k-means
init, EMA update, dead-code expiry, DDP buffer broadcast; core_vq.py:141-169, 217-229) are
out of scope for the synthesis path and raise."""
import typing as tp

import torch
from torch import nn

from ... import _lib as L
from ... import ops


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

    def _require_inference(self):
        if self.training:
            raise L.MttsError("VQ training (EMA / k-means / dead-code expiry) is outside the synthesis path")
        sig = (self.inited.data_ptr(), self.inited._version)
        if self._inited_sig != sig:     # one readback per (re)load, not per call
            if not bool(self.inited.detach().cpu().item()):
                raise L.MttsError("codebook not initialised (inited == 0): the reference would run k-means here "
                                  "(core_vq.py:141-149); load a trained codebook first")
            self._inited_sig = sig

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
        self._require_inference()
        shape = x.shape
        ind = self.quantize(x.reshape(-1, shape[-1])).view(*shape[:-1])
        return self.dequantize(ind), ind


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
        """eval-mode VectorQuantization.forward (core_vq.py:294-316): (quantize (B,D,N), ind (B,N), loss [0.])"""
        q, ind = self._codebook(ops.to_channels_last(x))
        loss = torch.zeros(1, device=x.device)
        return ops.to_channels_first(q), ind, loss


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
