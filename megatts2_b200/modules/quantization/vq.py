"""ResidualVectorQuantizer with the reference's surface (modules/quantization/vq.py:28-113)."""
import math
import typing as tp

import torch
from torch import nn

from .core_vq import ResidualVectorQuantization


class ResidualVectorQuantizer(nn.Module):
    def __init__(self, dimension: int = 256, n_q: int = 8, bins: int = 1024, decay: float = 0.99,
                 kmeans_init: bool = True, kmeans_iters: int = 50, threshold_ema_dead_code: int = 2):
        super().__init__()
        self.n_q = n_q
        self.dimension = dimension
        self.bins = bins
        self.decay = decay
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.vq = ResidualVectorQuantization(
            dim=dimension, codebook_size=bins, num_quantizers=n_q, decay=decay, kmeans_init=kmeans_init,
            kmeans_iters=kmeans_iters, threshold_ema_dead_code=threshold_ema_dead_code)

    def forward(self, x: torch.Tensor):
        """x (B, D, N) -> (quantized (B,D,N), codes (n_q,B,N) int64, losses (n_q,1))  (vq.py:69-81)."""
        return self.vq(x, n_q=self.n_q)

    def get_num_quantizers_for_bandwidth(self, frame_rate: int, bandwidth: tp.Optional[float] = None) -> int:
        bw_per_q = self.get_bandwidth_per_quantizer(frame_rate)
        n_q = self.n_q
        if bandwidth and bandwidth > 0.:
            n_q = int(max(1, math.floor(bandwidth * 1000 / bw_per_q)))
        return n_q

    def get_bandwidth_per_quantizer(self, frame_rate: int):
        return math.log2(self.bins) * frame_rate

    def encode(self, x: torch.Tensor, frame_rate: int = 0, bandwidth: tp.Optional[float] = None) -> torch.Tensor:
        return self.vq.encode(x, n_q=self.n_q)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes (n_q, B, N) int64 -> (B, D, N)  (vq.py:109-113)."""
        return self.vq.decode(codes)
